"""Training-step timing on one GPU: (a) frozen Base-16 encoder + trainable image tokenizer (the reference's AST / X-Ray
recipe), (b) full fine-tune.  Prints ms for forward and backward (CUDA events) and tokens/s.

    python tools/bench_train.py [B] [depth]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import metatransformer_b200 as mt  # noqa: E402


def run(B, depth, finetune):
    torch.manual_seed(0)
    enc = mt.build_encoder("base", depth=depth).cuda()
    enc = enc.train() if finetune else enc.eval().requires_grad_(False)
    tok = mt.Data2Seq("image", 768, token_dtype=torch.bfloat16).cuda()
    img = torch.rand(B, 3, 224, 224, device="cuda")

    def step():
        e = [torch.cuda.Event(True) for _ in range(3)]
        e[0].record()
        y = enc(tok(img))
        loss = y.float().square().mean()
        e[1].record()
        loss.backward()
        e[2].record()
        return e

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    fw = bw = 0.0
    iters = 5
    for _ in range(iters):
        e = step()
        torch.cuda.synchronize()
        fw += e[0].elapsed_time(e[1])
        bw += e[1].elapsed_time(e[2])
    fw, bw = fw / iters, bw / iters
    print(f"finetune={finetune} B={B} depth={depth}: forward {fw:.2f} ms, backward {bw:.2f} ms, "
          f"{B * 196 / (fw + bw) * 1e3 / 1e6:.3f} Mtok/s (fwd+bwd)", flush=True)


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    run(B, depth, finetune=False)
    run(B, depth, finetune=True)
