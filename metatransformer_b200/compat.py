"""Make the reference's import lines resolve to this package:

    import metatransformer_b200.compat as compat; compat.install()
    from timm.models.vision_transformer import Block          # README.md:116
    from Data2Seq import Data2Seq                             # README.md:117

`timm` is only shimmed when the real package is absent (the real one keeps working otherwise; then import
`metatransformer_b200.vision_transformer.Block` explicitly).
"""
import importlib
import importlib.util
import sys
import types


def _real_timm_present() -> bool:
    mod = sys.modules.get("timm")
    if mod is not None:
        return getattr(mod, "_mt_b200_shim", False) is False
    try:
        return importlib.util.find_spec("timm") is not None
    except (ValueError, ImportError):
        return False


def install(force_timm: bool = False) -> None:
    from . import data2seq, vision_transformer
    if force_timm or not _real_timm_present():
        timm = types.ModuleType("timm")
        timm._mt_b200_shim = True
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = vision_transformer.DropPath
        layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        # initialisers the reference forks import from timm.models.layers (Video/models/modeling_finetune.py:15,
        # Audio/src/models/ast_models.py:13): torch's own implementations
        import torch.nn as _nn
        layers.trunc_normal_ = _nn.init.trunc_normal_
        layers.lecun_normal_ = lambda t: _nn.init.trunc_normal_(t, std=(1.0 / max(1, t.shape[-1])) ** 0.5)
        timm.models = models
        models.vision_transformer = vision_transformer
        models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers,
                            "timm.models.vision_transformer": vision_transformer})
    # `from Data2Seq import Data2Seq` (reference scripts run with Data2Seq/ on sys.path, Data2Seq.py:9-15)
    sys.modules.setdefault("Data2Seq", importlib.import_module(data2seq.__name__ + ".Data2Seq"))
    for name in ("Image", "Video", "Acoustic", "Time_Series", "Hyper_Spectrum", "Text", "Graph"):
        sys.modules.setdefault(name, getattr(data2seq, name))
