"""How far can ANY fp16/bf16-forward gradient be from the fp32 gradient?  CPU study: the emulated 16-bit forward of
tools/emulate_precision.py with straight-through rounding and an EXACT fp32 autograd backward, against the fp32
oracle's autograd, on random-init weights (design study; not part of the product or the tests).
usage: python tools/emulate_grad_noise.py [n] [c] [pairs]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fgnn_oracle as O
import tools.emulate_precision as E

def ste(x, dt):
    return x + (x.to(dt).to(torch.float32) - x).detach()

def run(n, c, pairs, dt):
    gen = torch.Generator().manual_seed(0)
    sd = O.xavier_state_dict(2, c, 4, 3, gen, randomize_gn=True)
    data = [O.synthetic_pair(n, 0.2, 0.1, gen) for _ in range(pairs)]
    def loss_of(embed):
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        tot = 0.0
        for a, b in data:
            e1, e2 = embed(a, params), embed(b, params)
            s = e1.t() @ e2
            tot = tot + torch.nn.functional.cross_entropy(s, torch.arange(n), reduction="sum")
        loss = tot / (pairs * n)
        loss.backward()
        return float(loss), {k: v.grad for k, v in params.items()}
    l_ref, g_ref = loss_of(lambda x, p: O.node_embedding(x[None], p)[0])
    if dt is None:
        return
    E.rnd = lambda x, d: ste(x, d)
    l_em, g_em = loss_of(lambda x, p: E.emulate(x, p, dt))
    worst = 0
    errs = []
    for k in g_ref:
        if g_ref[k].abs().max() < 1e-6:
            continue
        e = float((g_em[k] - g_ref[k]).norm() / g_ref[k].norm())
        errs.append(e)
    errs.sort()
    print(f"n={n} c={c} {dt}: loss {l_em:.5f} vs {l_ref:.5f}; parameter-gradient rel err median {errs[len(errs)//2]:.3e} max {errs[-1]:.3e}")

if __name__ == "__main__":
    torch.set_num_threads(8)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    c = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    for dt in (torch.float16, torch.bfloat16):
        run(n, c, pairs, dt)
