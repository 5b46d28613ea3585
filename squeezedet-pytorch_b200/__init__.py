"""sqdet-b200: SqueezeDet's post-backbone detection path as hand-written sm_100a CUDA kernels
behind the reference's own class surface (PredictionResolver / SqueezeDet / SqueezeDetWithLoss /
Detector).  See DESIGN.md.  Heavy submodules are imported lazily so that ``synth`` (pure numpy)
stays importable without torch."""
__version__ = "0.1.0"

_LAZY = {
    "SqueezeDetBase": "model", "PredictionResolver": "model", "Loss": "model",
    "SqueezeDetWithLoss": "model", "SqueezeDet": "model", "Fire": "model",
    "Detector": "detector",
    "compute_deltas": "targets", "prepare_annotations": "targets", "generate_anchors": "targets",
    "kitti_config": "config", "make_config": "config",
    "lib": "_lib",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module(f"{__name__}.{_LAZY[name]}")
        return mod if name == "lib" else getattr(mod, name)
    raise AttributeError(name)
