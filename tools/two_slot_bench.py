"""Device-resident step throughput with the steps alternating over S streams (each with its own workspace slot and output
block) against the single-stream chain.  usage: python tools/two_slot_bench.py [batch] [steps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squeezedet_pytorch_b200 import ops, synth
dev = torch.device("cuda")
shp = synth.KITTI
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
R = 3
feats = [torch.relu(torch.randn((B, 768, *shp.grid_hw), device=dev)) for _ in range(R)]
w, b = synth.convdet_params(shp, 4321)
weight, bias = torch.from_numpy(w).to(dev), torch.from_numpy(b).to(dev)
packed = ops.pack_convdet_weights(weight)
anchors = torch.from_numpy(synth.anchor_table(shp).astype(np.float32)).to(dev)
for S in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(S)]
    dets = [ops._alloc_detections(B, shp.top_k, dev) for _ in range(S)]
    def step(i):
        with torch.cuda.stream(streams[i % S]):
            ops.head_detect(feats[i % R], weight, bias, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw,
                            shp.top_k, shp.nms_thresh, shp.score_thresh, packed=packed, out=dets[i % S], slot=i % S)
    for i in range(12):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    for i in range(K):
        step(i)
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"B={B} streams={S}: {ms * 1e3:.1f} us per step, {B / ms * 1e3:.0f} img/s")
    ref = ops.head_detect(feats[(K - 1) % R], weight, bias, anchors, shp.anchors_per_grid, shp.num_classes, shp.input_hw,
                          shp.top_k, shp.nms_thresh, shp.score_thresh, packed=packed)
    torch.cuda.synchronize()
    d = dets[(K - 1) % S]
    assert torch.equal(ref.count, d.count) and torch.equal(ref.anchor, d.anchor), "slot result differs"
