"""Generate golden vectors by running the REFERENCE's own code (build container only: needs /root/reference).

    python tests/golden/make_golden.py

What is executed from the reference tree (never copied into this repo):
  * PointCloud/openpoints/models/layers/attention.py  -> classes Attention, Block   ("borrowed from TIMM", :2)
  * PointCloud/openpoints/models/layers/mlp.py        -> class Mlp
    Their package-relative imports (`from . import create_norm, create_act, DropPath, ...`, `from .helpers import
    to_2tuple`) pull in openpoints' registry machinery, which needs the absent `multimethod` package; we satisfy
    exactly those names with tiny stand-ins (nn.LayerNorm(eps=1e-5) — the README path, README.md:131 — and nn.GELU())
    and exec the two source files unmodified.
  * Data2Seq/Image.py       -> PatchEmbed (imports only torch)
  * Data2Seq/Time_Series.py -> DataEmbedding (imports only torch)
  * Audio/src/models/ast_models.py        -> class PatchEmbed (:18-33) with its `proj` replaced exactly as ASTModel does
    (:86: Conv2d(1, d, (16,16), stride=(fstride, tstride)))
  * Video/models/modeling_finetune.py     -> class PatchEmbed (:263-297), the tubelet Conv3d tokenizer
    Both files `import timm` / `wget` at the top; those modules are absent here, so tiny stand-ins providing only
    `to_2tuple`, `trunc_normal_`, `drop_path`, `register_model` are placed in sys.modules and the files are exec'd
    unmodified (only the two PatchEmbed classes are used).
The vectors land in tests/golden/*.npz (small, committed).  /root/reference does not exist on the GPU box, so tests
only ever read the .npz files.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _load_reference_block_classes():
    layers = REF / "PointCloud/openpoints/models/layers"
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = [str(layers)]

    def create_norm(norm_args, channels, dimension=None):
        return nn.LayerNorm(channels)  # README.md:131 passes a bare nn.LayerNorm -> eps 1e-5

    def create_act(act_args):
        return nn.GELU()               # README.md:132 act_layer=nn.GELU

    class DropPath(nn.Identity):       # only constructed when drop_path > 0 (attention.py:49); never here
        def __init__(self, p=0.0):
            super().__init__()

    pkg.create_norm = create_norm
    pkg.create_act = create_act
    pkg.DropPath = DropPath
    pkg.trunc_normal_ = nn.init.trunc_normal_
    pkg.lecun_normal_ = lambda t: t
    sys.modules["refpkg"] = pkg
    helpers = types.ModuleType("refpkg.helpers")
    helpers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    sys.modules["refpkg.helpers"] = helpers

    def load(name, fname):
        spec = importlib.util.spec_from_file_location(f"refpkg.{name}", layers / fname)
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "refpkg"
        sys.modules[f"refpkg.{name}"] = mod
        spec.loader.exec_module(mod)
        return mod

    mlp_mod = load("mlp", "mlp.py")
    pkg.Mlp = mlp_mod.Mlp
    att_mod = load("attention", "attention.py")
    return att_mod.Attention, att_mod.Block, mlp_mod.Mlp


def _load_file_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stub_timm():
    """Stand-ins for the names the two model files import from timm / wget at module level."""
    if "timm" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    registry = types.ModuleType("timm.models.registry")
    layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
    layers.trunc_normal_ = nn.init.trunc_normal_
    layers.drop_path = lambda x, drop_prob=0.0, training=False: x
    registry.register_model = lambda fn: fn
    timm.models, models.layers, models.registry = models, layers, registry
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers,
                        "timm.models.registry": registry, "wget": types.ModuleType("wget")})


def _randomise(module: nn.Module, seed: int):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    Attention, Block, Mlp = _load_reference_block_classes()

    # ---- encoder: nn.Sequential of the reference's Block, two configs (head_dim 64 like Base/Large, and 32) ----
    for tag, (dim, heads, depth, B, N) in {"enc_d128_h2_L2": (128, 2, 2, 2, 37), "enc_d96_h3_L2": (96, 3, 2, 3, 50)}.items():
        enc = nn.Sequential(*[Block(dim=dim, num_heads=heads, mlp_ratio=4., qkv_bias=True) for _ in range(depth)]).eval()
        _randomise(enc, seed=len(tag) + dim)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(B, N, dim, generator=g)
        y = enc(x)
        # per-stage intermediates of block 0 for finer-grained checks
        b0 = enc[0]
        h1 = b0.norm1(x)
        a1 = b0.attn(h1)
        x1 = x + a1
        m1 = b0.mlp(b0.norm2(x1))
        arrays = {"x": x.numpy(), "y": y.numpy(), "b0_norm1": h1.numpy(), "b0_attn": a1.numpy(), "b0_mlp": m1.numpy(),
                  "meta": np.array([dim, heads, depth, B, N])}
        for k, v in enc.state_dict().items():
            arrays["sd." + k] = v.numpy()
        np.savez_compressed(OUT / f"{tag}.npz", **arrays)
        print(tag, "y abs max", float(y.abs().max()))

    # ---- Data2Seq/Image.py PatchEmbed (the 'image' / 'infrared' / 'x-ray' tokenizer) ----
    image_mod = _load_file_module("ref_image", REF / "Data2Seq/Image.py")
    pe = image_mod.PatchEmbed(embed_dim=64).eval()
    _randomise(pe, seed=3)
    g = torch.Generator().manual_seed(11)
    x = torch.randint(-8, 9, (2, 3, 224, 224), generator=g).float() / 8.0   # exactly representable, compresses well
    y = pe(x)
    np.savez_compressed(OUT / "image_patch_embed.npz", x=x.numpy().astype(np.float16), y=y.numpy(),
                        w=pe.proj.weight.numpy(), b=pe.proj.bias.numpy())
    print("image_patch_embed", tuple(y.shape))

    # ---- Data2Seq/Time_Series.py DataEmbedding (works with c_in=; eval => dropout identity) ----
    ts_mod = _load_file_module("ref_ts", REF / "Data2Seq/Time_Series.py")
    de = ts_mod.DataEmbedding(c_in=7, d_model=64).eval()
    g = torch.Generator().manual_seed(13)
    x = torch.randn(3, 96, 7, generator=g)
    y = de(x)
    np.savez_compressed(OUT / "time_series_embed.npz", x=x.numpy(), y=y.numpy(),
                        w=de.value_embedding.tokenConv.weight.numpy())
    print("time_series_embed", tuple(y.shape))
    # with calendar features: fixed sinusoid tables (freq 't': month, day, weekday, hour, minute) and the timeF Linear
    g = torch.Generator().manual_seed(14)
    de_f = ts_mod.DataEmbedding(c_in=7, d_model=64, embed_type='fixed', freq='t').eval()
    xm = torch.stack([torch.randint(0, 13, (3, 96), generator=g), torch.randint(0, 32, (3, 96), generator=g),
                      torch.randint(0, 7, (3, 96), generator=g), torch.randint(0, 24, (3, 96), generator=g),
                      torch.randint(0, 4, (3, 96), generator=g)], dim=-1).float()
    de_t = ts_mod.DataEmbedding(c_in=7, d_model=64, embed_type='timeF', freq='h').eval()
    xt = torch.rand(3, 96, 4, generator=g) - 0.5
    np.savez_compressed(OUT / "time_series_embed_mark.npz", x=x.numpy(), x_mark_fixed=xm.numpy(), y_fixed=de_f(x, xm).detach().numpy(),
                        w_fixed=de_f.value_embedding.tokenConv.weight.detach().numpy(), x_mark_timef=xt.numpy(),
                        y_timef=de_t(x, xt).detach().numpy(), w_timef=de_t.value_embedding.tokenConv.weight.detach().numpy(),
                        w_timef_embed=de_t.temporal_embedding.embed.weight.detach().numpy())
    print("time_series_embed_mark")

    # ---- Data2Seq/Graph.py GraphFeatureTokenizer (imports only torch): 3 ragged graphs, every identifier encoder on ----
    gr_mod = _load_file_module("ref_graph", REF / "Data2Seq/Graph.py")
    torch.manual_seed(21)
    gt = gr_mod.GraphFeatureTokenizer(num_atoms=40, num_edges=12, rand_node_id=True, rand_node_id_dim=16, orf_node_id=True,
                                      orf_node_id_dim=16, lap_node_id=True, lap_node_id_k=8, lap_node_id_sign_flip=False,
                                      lap_node_id_eig_dropout=0.0, type_id=True, hidden_dim=64, n_layers=2).eval()
    g = torch.Generator().manual_seed(23)
    node_num, edge_num = [5, 3, 6], [8, 4, 10]
    n_tot, e_tot = sum(node_num), sum(edge_num)
    node_data = torch.randint(1, 40, (n_tot, 3), generator=g)
    edge_data = torch.randint(1, 12, (e_tot, 2), generator=g)
    edge_index = torch.cat([torch.randint(0, n, (2, e), generator=g) for n, e in zip(node_num, edge_num)], dim=1)
    lap_eigvec = torch.randn(n_tot, 8, generator=g)
    batch = dict(node_data=node_data, in_degree=torch.zeros(n_tot, dtype=torch.long), out_degree=torch.zeros(n_tot, dtype=torch.long),
                 node_num=node_num, lap_eigvec=lap_eigvec, lap_eigval=torch.zeros(n_tot, 8), edge_index=edge_index,
                 edge_data=edge_data, edge_num=edge_num)
    torch.manual_seed(5)
    with torch.no_grad():
        feat, pmask, pidx = gt(batch)
    # the two random identifier draws of that call, replayed in the same order (Graph.py:264-281)
    torch.manual_seed(5)
    rid = torch.nn.functional.normalize(torch.rand(n_tot, 16), p=2, dim=1)
    node_mask = gt.get_node_mask(node_num, torch.device("cpu"))
    orf = gr_mod.gaussian_orthogonal_random_matrix_batched(3, max(node_num), max(node_num))[node_mask]
    orf = torch.nn.functional.normalize(torch.nn.functional.pad(orf, (0, 16 - max(node_num)), value=0.0), p=2, dim=1)
    np.savez_compressed(OUT / "graph_tokenizer.npz", node_data=node_data.numpy(), edge_data=edge_data.numpy(),
                        edge_index=edge_index.numpy(), lap_eigvec=lap_eigvec.numpy(), node_num=np.array(node_num),
                        edge_num=np.array(edge_num), rand_node_id=rid.numpy(), orf_node_id=orf.numpy(), y=feat.numpy(),
                        padding_mask=pmask.numpy(), padded_index=pidx.numpy(),
                        **{"sd." + k: v.numpy() for k, v in gt.state_dict().items()})
    print("graph_tokenizer", tuple(feat.shape))

    # ---- AST tokenizer: the reference's PatchEmbed with the overlapping-stride conv ASTModel installs (:86) ----
    _stub_timm()
    ast_mod = _load_file_module("ref_ast_models", REF / "Audio/src/models/ast_models.py")
    ape = ast_mod.PatchEmbed(img_size=224, patch_size=16, in_chans=1, embed_dim=64).eval()
    ape.proj = nn.Conv2d(1, 64, kernel_size=(16, 16), stride=(10, 10))      # ast_models.py:86 (fstride = tstride = 10)
    _randomise(ape, seed=5)
    g = torch.Generator().manual_seed(17)
    x = torch.randint(-8, 9, (2, 1, 64, 100), generator=g).float() / 8.0     # (B, 1, F, T) as built at :152-153
    y = ape(x)
    np.savez_compressed(OUT / "audio_patch_embed.npz", x=x.numpy().astype(np.float16), y=y.numpy(),
                        w=ape.proj.weight.numpy(), b=ape.proj.bias.numpy(), stride=np.array([10, 10]))
    print("audio_patch_embed", tuple(y.shape))

    # ---- VideoMAE tubelet tokenizer ----
    vid_mod = _load_file_module("ref_modeling_finetune", REF / "Video/models/modeling_finetune.py")
    vpe = vid_mod.PatchEmbed(img_size=32, patch_size=16, in_chans=3, embed_dim=64, num_frames=4, tubelet_size=2).eval()
    _randomise(vpe, seed=6)
    g = torch.Generator().manual_seed(19)
    x = torch.randint(-8, 9, (2, 3, 4, 32, 32), generator=g).float() / 8.0
    y = vpe(x)
    np.savez_compressed(OUT / "video_patch_embed.npz", x=x.numpy().astype(np.float16), y=y.numpy(),
                        w=vpe.proj.weight.numpy(), b=vpe.proj.bias.numpy())
    print("video_patch_embed", tuple(y.shape))


if __name__ == "__main__":
    main()
