"""CPU oracle for Meta-Transformer's shared encoder path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg (and bench.py's parity
spot check, as the checker) may import this module;
the product package (metatransformer_b200/) never does — it fails loudly when the CUDA extension is missing.

What is restated (all arithmetic in plain torch CPU ops written out explicitly, float32 or float64):
  * timm 0.4.12 `Mlp` / `Attention` / `Block` — timm is a pip dependency of the reference that is NOT vendored
    (pins: X-Ray/requirements.txt:9, Image/environment.yaml:77, Video/requirements.txt:13 = 0.4.12;
    Audio/requirements.txt:8 = 0.4.5; Tabular/README.md:31 = 0.8.0.dev0).  The restatement follows the verbatim
    in-tree copy: PointCloud/openpoints/models/layers/attention.py:12-38 (Attention), :41-58 (Block),
    PointCloud/openpoints/models/layers/mlp.py:11-35 (Mlp); cross-read Image/detection/mmdet_custom/models/
    backbones/base/vit.py:90-115.  Hyper-parameters from README.md:125-147 (LayerNorm eps = nn.LayerNorm default 1e-5,
    act = nn.GELU() = exact erf, qkv_bias=True, mlp_ratio=4).
  * Data2Seq conv tokenizers: Data2Seq/Image.py:4-28; Audio/src/models/ast_models.py:18-33,86,152-153;
    Video/models/modeling_finetune.py:263-297; Data2Seq/Time_Series.py:8-42,109-126.

PARITY PINNING.  The reference holds NO golden vector / known-answer test for this path, and
timm itself cannot be imported here.  The oracle is therefore pinned against outputs of the reference's OWN code
run in the build container: tests/golden/make_golden.py executes the reference's in-tree Attention/Block/Mlp
classes (attention.py / mlp.py loaded from /root/reference with their package-relative imports stubbed),
Data2Seq/Image.py, Data2Seq/Time_Series.py and — with stand-ins for the `timm` / `wget` names their files import at
module level — the PatchEmbed classes of Audio/src/models/ast_models.py and Video/models/modeling_finetune.py, on seeded
inputs, and stores the vectors in tests/golden/*.npz; tests/test_oracle_golden.py checks this file against them.
Also allowed to call this module, as a checker only: the two-sequence parity figure bench.py prints beside its
throughput.  Status: every function below is pinned against reference code run here.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------------------
# elementary ops, written out (no torch.nn.functional composite calls)
# --------------------------------------------------------------------------------------------------------------
def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm(d): biased variance over the last dim (attention.py:46,50 via create_norm; README.md:131)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """nn.Linear: y = x W^T + b, W stored [out, in]."""
    y = x @ weight.t()
    return y if bias is None else y + bias


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default (approximate='none'): 0.5 x (1 + erf(x / sqrt 2))  (README.md:132)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


# --------------------------------------------------------------------------------------------------------------
# timm Block
# --------------------------------------------------------------------------------------------------------------
def attention(x: Tensor, qkv_w: Tensor, qkv_b: Optional[Tensor], proj_w: Tensor, proj_b: Tensor, num_heads: int,
              qk_scale: Optional[float] = None) -> Tensor:
    """Attention.forward, attention.py:26-38 (dropouts are identity at p=0 / eval)."""
    B, N, C = x.shape
    dh = C // num_heads
    scale = qk_scale if qk_scale is not None else dh ** -0.5              # attention.py:18-19
    qkv = linear(x, qkv_w, qkv_b).reshape(B, N, 3, num_heads, dh).permute(2, 0, 3, 1, 4)   # :28
    q, k, v = qkv[0], qkv[1], qkv[2]                                      # :29
    attn = (q @ k.transpose(-2, -1)) * scale                              # :31
    attn = attn - attn.max(dim=-1, keepdim=True).values                   # softmax, written out (:32)
    attn = attn.exp()
    attn = attn / attn.sum(dim=-1, keepdim=True)
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)                     # :35
    return linear(out, proj_w, proj_b)                                    # :36


def mlp(x: Tensor, fc1_w: Tensor, fc1_b: Tensor, fc2_w: Tensor, fc2_b: Tensor) -> Tensor:
    """Mlp.forward, mlp.py:28-35."""
    return linear(gelu_erf(linear(x, fc1_w, fc1_b)), fc2_w, fc2_b)


def block(x: Tensor, sd: Dict[str, Tensor], prefix: str, num_heads: int, eps: float = 1e-5) -> Tensor:
    """Block.forward, attention.py:55-58: x = x + attn(norm1(x)); x = x + mlp(norm2(x)) (drop_path = Identity)."""
    g = lambda k: sd[prefix + k]  # noqa: E731
    h = layer_norm(x, g("norm1.weight"), g("norm1.bias"), eps)
    x = x + attention(h, g("attn.qkv.weight"), sd.get(prefix + "attn.qkv.bias"), g("attn.proj.weight"),
                      g("attn.proj.bias"), num_heads)
    h = layer_norm(x, g("norm2.weight"), g("norm2.bias"), eps)
    x = x + mlp(h, g("mlp.fc1.weight"), g("mlp.fc1.bias"), g("mlp.fc2.weight"), g("mlp.fc2.bias"))
    return x


def encoder(x: Tensor, sd: Dict[str, Tensor], depth: int, num_heads: int, eps: float = 1e-5) -> Tensor:
    """nn.Sequential(*[Block(...)]*depth)(x) with the flat state-dict keys '{i}.norm1.weight', ... (README.md:125-149)."""
    for i in range(depth):
        x = block(x, sd, f"{i}.", num_heads, eps)
    return x


# --------------------------------------------------------------------------------------------------------------
# Data2Seq tokenizers
# --------------------------------------------------------------------------------------------------------------
def _im2row(x: Tensor, kernel: Sequence[int], stride: Sequence[int]) -> Tensor:
    """x [B,C,T,H,W] -> [B, Tp*Hp*Wp, C*kt*kh*kw]; K order (c,kt,kh,kw) = weight.view(d,-1); tokens (t,h,w) row-major."""
    B, C = x.shape[:2]
    p = x.unfold(2, kernel[0], stride[0]).unfold(3, kernel[1], stride[1]).unfold(4, kernel[2], stride[2])
    Tp, Hp, Wp = p.shape[2:5]
    return p.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(B, Tp * Hp * Wp, C * kernel[0] * kernel[1] * kernel[2])


def patch_embed_2d(x: Tensor, weight: Tensor, bias: Optional[Tensor], stride: Sequence[int]) -> Tensor:
    """Image.PatchEmbed.forward (Data2Seq/Image.py:19-28): Conv2d(k, stride) -> flatten(2).transpose(1,2); no norm
    (:27 is commented out).  Cross-correlation, weight [d, C, kh, kw].  Also the AST spectrogram conv
    (Audio/src/models/ast_models.py:86, 29-33)."""
    d, C, kh, kw = weight.shape
    rows = _im2row(x.unsqueeze(2), (1, kh, kw), (1, stride[0], stride[1]))
    return linear(rows, weight.reshape(d, -1), bias)


def patch_embed_3d(x: Tensor, weight: Tensor, bias: Optional[Tensor], stride: Sequence[int]) -> Tensor:
    """VideoMAE PatchEmbed.forward (Video/models/modeling_finetune.py:285-297): Conv3d(tubelet) -> flatten/transpose."""
    d, C, kt, kh, kw = weight.shape
    rows = _im2row(x, (kt, kh, kw), stride)
    return linear(rows, weight.reshape(d, -1), bias)


def audio_input_prep(fbank: Tensor) -> Tensor:
    """(B, T, F) -> (B, 1, F, T)  (Audio/src/models/ast_models.py:152-153)."""
    return fbank.unsqueeze(1).transpose(2, 3)


def sinusoid_pe(length: int, d_model: int, dtype=torch.float32) -> Tensor:
    """PositionalEmbedding (Data2Seq/Time_Series.py:8-26)."""
    pe = torch.zeros(length, d_model, dtype=torch.float32)
    position = torch.arange(0, length).float().unsqueeze(1)
    div_term = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(dtype)


def time_series_embed(x: Tensor, conv_w: Tensor) -> Tensor:
    """DataEmbedding.forward with x_mark=None, eval mode (Data2Seq/Time_Series.py:120-126, 29-42):
    Conv1d(c_in, d, k=3, padding=1, circular, bias=False) over time + sinusoid PE.  x [B,T,c_in], conv_w [d,c_in,3]."""
    B, T, Cin = x.shape
    xp = torch.cat([x[:, -1:, :], x, x[:, :1, :]], dim=1)            # circular padding of 1 on the time axis
    rows = torch.stack([xp[:, 0:T], xp[:, 1:T + 1], xp[:, 2:T + 2]], dim=-1)   # [B,T,c_in,3]
    tok = rows.reshape(B, T, Cin * 3) @ conv_w.reshape(conv_w.shape[0], -1).t()
    return tok + sinusoid_pe(T, conv_w.shape[0], x.dtype)


def graph_tokens(sd: Dict[str, Tensor], node_data: Tensor, edge_data: Tensor, edge_index: Tensor, node_num: Sequence[int],
                 edge_num: Sequence[int], lap_eigvec: Optional[Tensor] = None, rand_node_id: Optional[Tensor] = None,
                 orf_node_id: Optional[Tensor] = None) -> Tensor:
    """GraphFeatureTokenizer.forward in eval mode (Data2Seq/Graph.py:234-305) with the random node identifiers given as
    inputs: per token  sum_f atom|edge embedding (:254-255) + rand / orf / lap encoders on [id_u ; id_v] (:263-294) +
    order embedding of (u == v) (:296-297); sequence [graph_token, null_token, nodes, edges, zero padding] (:299-304).
    sd = the tokenizer's state-dict.  Written per graph with plain indexing (no batched scatter)."""
    D = sd["atom_encoder.weight"].shape[1]
    b, T = len(node_num), max(n + e for n, e in zip(node_num, edge_num))
    out = torch.zeros(b, T + 2, D, dtype=sd["atom_encoder.weight"].dtype)
    n0 = e0 = 0
    for g in range(b):
        n, e = node_num[g], edge_num[g]
        u = torch.cat([torch.arange(n), edge_index[0, e0:e0 + e]])
        v = torch.cat([torch.arange(n), edge_index[1, e0:e0 + e]])
        feat = torch.cat([sd["atom_encoder.weight"][node_data[n0:n0 + n]].sum(-2),
                          sd["edge_encoder.weight"][edge_data[e0:e0 + e]].sum(-2)])
        for name, ids in (("rand_encoder", rand_node_id), ("orf_encoder", orf_node_id), ("lap_encoder", lap_eigvec)):
            if ids is not None and name + ".weight" in sd:
                loc = ids[n0:n0 + n].to(feat.dtype)
                k = sd[name + ".weight"].shape[1] // 2
                loc = loc[:, :k] if loc.shape[1] >= k else torch.nn.functional.pad(loc, (0, k - loc.shape[1]))
                feat = feat + torch.cat([loc[u], loc[v]], dim=1) @ sd[name + ".weight"].t()
        if "order_encoder.weight" in sd:
            feat = feat + sd["order_encoder.weight"][(u == v).long()]
        out[g, 0], out[g, 1] = sd["graph_token.weight"][0], sd["null_token.weight"][0]
        out[g, 2:2 + n + e] = feat
        n0, e0 = n0 + n, e0 + e
    return out


# --------------------------------------------------------------------------------------------------------------
# seeded synthetic weights (the pretrained checkpoints are external links, README.md:103-104)
# --------------------------------------------------------------------------------------------------------------
def make_encoder_state_dict(dim: int, depth: int, mlp_ratio: float = 4.0, seed: int = 0, qkv_bias: bool = True,
                            dtype=torch.float32) -> Dict[str, Tensor]:
    """Flat state-dict of nn.Sequential(Block...).  trunc_normal_(std=.02)-like weights,
    N(0,.02) biases, LN gamma 1+N(0,.02), beta N(0,.02) so that every bias / affine path is exercised."""
    g = torch.Generator().manual_seed(seed)
    hidden = int(dim * mlp_ratio)
    sd: Dict[str, Tensor] = {}

    def w(*shape):
        return (torch.randn(*shape, generator=g) * 0.02).clamp_(-0.04, 0.04).to(dtype)

    def b(n):
        return (torch.randn(n, generator=g) * 0.02).to(dtype)

    for i in range(depth):
        p = f"{i}."
        sd[p + "norm1.weight"] = (1.0 + torch.randn(dim, generator=g) * 0.02).to(dtype)
        sd[p + "norm1.bias"] = b(dim)
        sd[p + "attn.qkv.weight"] = w(3 * dim, dim)
        if qkv_bias:
            sd[p + "attn.qkv.bias"] = b(3 * dim)
        sd[p + "attn.proj.weight"] = w(dim, dim)
        sd[p + "attn.proj.bias"] = b(dim)
        sd[p + "norm2.weight"] = (1.0 + torch.randn(dim, generator=g) * 0.02).to(dtype)
        sd[p + "norm2.bias"] = b(dim)
        sd[p + "mlp.fc1.weight"] = w(hidden, dim)
        sd[p + "mlp.fc1.bias"] = b(hidden)
        sd[p + "mlp.fc2.weight"] = w(dim, hidden)
        sd[p + "mlp.fc2.bias"] = b(dim)
    return sd


def make_tokens(B: int, N: int, dim: int, seed: int = 1234, dtype=torch.float32) -> Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, N, dim, generator=g).to(dtype)


def encoder_flops(B: int, N: int, dim: int, depth: int, mlp_ratio: float = 4.0) -> float:
    """Algorithmic forward FLOPs B*N*L*(24 d^2 + 4 N d) for mlp_ratio 4."""
    hidden = dim * mlp_ratio
    per_token = 2 * (3 * dim * dim + dim * dim + 2 * dim * hidden) + 4 * N * dim
    return float(B) * N * depth * per_token
