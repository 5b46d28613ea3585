"""Test / measurement infrastructure: the reference's hot path restated OP FOR OP with torch and torchvision calls, so that
it can run on any torch device -- in particular on the B200 itself (the reference's own GPU path: cuDNN conv + ~40 ATen
kernels + a per-image Python filter loop with torchvision.ops.nms and its host syncs).  The reference's code cannot
travel to the GPU box; this file follows it line by line instead:

  SqueezeDetBase.forward tail      src/model/squeezedet.py:83-87
  PredictionResolver.forward       src/model/squeezedet.py:109-120, safe_softmax / deltas_to_boxes src/model/modules.py:17-45,66-68
  SqueezeDet.forward               src/model/squeezedet.py:197-206
  Detector.filter                  src/engine/detector.py:87-122

Used by bench.py for the `gpu_reference_baseline` figure only; nothing in the product imports it."""
import torch
import torchvision


def xywh_to_xyxy(boxes_xywh):                                   # modules.py:17-24
    return torch.cat([boxes_xywh[..., [0]] - 0.5 * (boxes_xywh[..., [2]] - 1), boxes_xywh[..., [1]] - 0.5 * (boxes_xywh[..., [3]] - 1),
                      boxes_xywh[..., [0]] + 0.5 * (boxes_xywh[..., [2]] - 1), boxes_xywh[..., [1]] + 0.5 * (boxes_xywh[..., [3]] - 1)], dim=-1)


def deltas_to_boxes(deltas, anchors, input_size):               # modules.py:27-45
    boxes_xywh = torch.cat([anchors[..., [0]] + anchors[..., [2]] * deltas[..., [0]],
                            anchors[..., [1]] + anchors[..., [3]] * deltas[..., [1]],
                            anchors[..., [2]] * torch.exp(deltas[..., [2]]),
                            anchors[..., [3]] * torch.exp(deltas[..., [3]])], dim=2)
    boxes_xyxy = xywh_to_xyxy(boxes_xywh)
    boxes_xyxy[..., [0, 2]] = torch.clamp(boxes_xyxy[..., [0, 2]], 0, input_size[1] - 1)
    boxes_xyxy[..., [1, 3]] = torch.clamp(boxes_xyxy[..., [1, 3]], 0, input_size[0] - 1)
    return boxes_xyxy


def safe_softmax(probs, dim=None):                              # modules.py:66-68
    exp = torch.exp(probs - torch.max(probs, dim=dim, keepdim=True)[0])
    return exp / torch.sum(exp, dim=dim, keepdim=True)


def head_forward(feat, weight, bias, anchors, num_classes, num_anchors, input_size):
    """features -> {'class_ids', 'scores', 'boxes'} like SqueezeDet.forward after the backbone."""
    x = torch.nn.functional.conv2d(feat, weight, bias, stride=1, padding=1)        # squeezedet.py:83
    x = x.permute(0, 2, 3, 1).contiguous()                                       # :85
    pred = x.view(-1, num_anchors, num_classes + 5)                              # :86
    pred_class_probs = safe_softmax(pred[..., :num_classes].contiguous(), dim=-1)  # :110
    pred_scores = torch.sigmoid(pred[..., num_classes:num_classes + 1].contiguous())  # :114
    pred_deltas = pred[..., num_classes + 1:].contiguous()                       # :115
    pred_boxes = deltas_to_boxes(pred_deltas, anchors.to(pred.device), input_size)  # :116-118 (re-uploaded every call)
    pred_class_probs *= pred_scores                                              # :200
    pred_class_ids = torch.argmax(pred_class_probs, dim=2)                       # :201
    pred_scores = torch.max(pred_class_probs, dim=2)[0]                          # :202
    return {"class_ids": pred_class_ids, "scores": pred_scores, "boxes": pred_boxes}


def filter_image(det, num_classes, keep_top_k, nms_thresh, score_thresh):         # detector.py:87-122
    orders = torch.argsort(det["scores"], descending=True)[:keep_top_k]
    class_ids = det["class_ids"][orders]
    scores = det["scores"][orders]
    boxes = det["boxes"][orders, :]
    filtered_class_ids, filtered_scores, filtered_boxes = [], [], []
    for cls_id in range(num_classes):
        idx_cur_class = (class_ids == cls_id)
        if torch.sum(idx_cur_class) == 0:                       # host sync, like the reference
            continue
        class_ids_cur_class = class_ids[idx_cur_class]
        scores_cur_class = scores[idx_cur_class]
        boxes_cur_class = boxes[idx_cur_class, :]
        keeps = torchvision.ops.nms(boxes_cur_class, scores_cur_class, nms_thresh)
        filtered_class_ids.append(class_ids_cur_class[keeps])
        filtered_scores.append(scores_cur_class[keeps])
        filtered_boxes.append(boxes_cur_class[keeps, :])
    filtered_class_ids = torch.cat(filtered_class_ids)
    filtered_scores = torch.cat(filtered_scores)
    filtered_boxes = torch.cat(filtered_boxes, dim=0)
    keeps = filtered_scores > score_thresh
    if torch.sum(keeps) == 0:                                   # host sync, like the reference
        return None
    return {"class_ids": filtered_class_ids[keeps], "scores": filtered_scores[keeps], "boxes": filtered_boxes[keeps, :]}


@torch.no_grad()
def detect(feat, weight, bias, anchors_cpu, num_classes, input_size, keep_top_k, nms_thresh, score_thresh):
    """One batch: Detector.detect's device work (detector.py:20-37) -- dense forward, then the per-image filter loop
    with its .cpu() transfers."""
    dets = head_forward(feat, weight, bias, anchors_cpu, num_classes, anchors_cpu.shape[1], input_size)
    out = []
    for b in range(dets["class_ids"].shape[0]):
        det = filter_image({k: v[b] for k, v in dets.items()}, num_classes, keep_top_k, nms_thresh, score_thresh)
        out.append(None if det is None else {k: v.cpu().numpy() for k, v in det.items()})
    return out
