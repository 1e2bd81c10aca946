"""Emulate the numerics of the planned low-precision pipeline on CPU (design study).
Compares node embeddings against the fp64 oracle.  Not part of the product or the tests."""
import sys, os, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fgnn_oracle as O

def rnd(x, dt):
    return x.to(dt).to(torch.float32)

def emulate(x, sd, dt, center=False, normalized_store=False, root="node_embedder.ne_bm_block", dt_inner=None, dt_w1=None):
    """x: (C0,n,n) fp32 one graph."""
    dti = dt if dt_inner is None else dt_inner        # weights and hidden activations (planes stay in dt)
    nb, depth = O.count_blocks(sd, root)
    n = x.shape[-1]
    P = n * n
    xs = rnd(x.reshape(x.shape[0], P), dt)          # stored block input (pre-norm)
    a_prev = torch.ones(x.shape[0]); s_prev = torch.zeros(x.shape[0])
    def mlp(inp_list, pre):
        # inp_list: list of (stored (Ci,P), a (Ci), s (Ci)); returns stored pre-norm yhat (bf16), a, s
        W1 = sd[f"{pre}.convs.0.weight"].reshape(sd[f"{pre}.convs.0.weight"].shape[0], -1)
        b = sd[f"{pre}.convs.0.bias"].clone()
        acc = 0; off = 0
        for st, a, s in inp_list:
            ci = st.shape[0]
            Wp = W1[:, off:off+ci]
            acc = acc + rnd(Wp * a[None, :], dti if dt_w1 is None else dt_w1) @ st
            b = b + Wp @ s
            off += ci
        h = acc + b[:, None]
        for k in range(1, depth):
            h = rnd(torch.relu(h), dti)
            Wk = sd[f"{pre}.convs.{k}.weight"].reshape(h.shape[0], -1)
            h = rnd(Wk, dti) @ h
            if k < depth - 1:
                h = h + sd[f"{pre}.convs.{k}.bias"][:, None]
        # h = last conv output without bias (bias cancels in GraphNorm)
        mu = h.mean(1); var = (h * h).mean(1) - mu * mu
        gw = sd[f"{pre}.gn.weight"].reshape(-1); gb = sd[f"{pre}.gn.bias"].reshape(-1)
        a = gw / (2 * torch.sqrt(n * (var + 1e-5)))
        if normalized_store:
            y = rnd(a[:, None] * (h - mu[:, None]) + gb[:, None], dt)
            return y, torch.ones_like(a), torch.zeros_like(a)
        if center:
            idx = torch.arange(0, P, 61)
            mu_est = h[:, idx].mean(1)
            st = rnd(h - mu_est[:, None], dt)
            s = gb - a * (mu - mu_est)
        else:
            st = rnd(h, dt)
            s = gb - a * mu
        return st, a, s
    for i in range(1, nb + 1):
        pre = f"{root}{i}"
        y1, a1, s1 = mlp([(xs, a_prev, s_prev)], pre + "_mlp1")
        y2, a2, s2 = mlp([(xs, a_prev, s_prev)], pre + "_mlp2")
        C = y1.shape[0]
        Y1 = y1.reshape(C, n, n); Y2 = y2.reshape(C, n, n)
        D = torch.matmul(Y1, Y2)
        r1 = Y1.sum(2); c2 = Y2.sum(1)
        mult = (a1 * a2)[:, None, None] * D + (a1 * s2)[:, None, None] * r1[:, :, None] \
            + (s1 * a2)[:, None, None] * c2[:, None, :] + (s1 * s2 * n)[:, None, None]
        mult = rnd(mult.reshape(C, P), dt)
        y3, a3, s3 = mlp([(mult, torch.ones(C), torch.zeros(C)), (xs, a_prev, s_prev)], pre + "_mlp3")
        xs, a_prev, s_prev = y3, a3, s3
    out = a_prev[:, None, None] * xs.reshape(-1, n, n) + s_prev[:, None, None]
    return out.max(-1)[0]

if __name__ == "__main__":
    torch.set_num_threads(8)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    c = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    reg = len(sys.argv) > 3
    for seed in (0, 1):
        gen = torch.Generator().manual_seed(seed)
        sd = O.xavier_state_dict(2, c, 4, 3, gen)
        x, _ = O.synthetic_pair(n, 0.2, 0.1, gen, regular_degree=int(0.2 * n) if reg else None)
        sd64 = {k: v.double() for k, v in sd.items()}
        ref = O.node_embedding(x[None].double(), sd64)[0]
        ref32 = O.node_embedding(x[None], sd)[0]
        def err(e): return float((e.double() - ref).norm() / ref.norm())
        print(f"n={n} c={c} seed={seed} fp32-oracle-vs-fp64: {err(ref32):.2e}")
        for dt in (torch.bfloat16, torch.float16):
            for kw in (dict(normalized_store=True), dict(), dict(center=True)):
                t = time.time()
                e = emulate(x, sd, dt, **kw)
                print(f"   {str(dt):16s} {str(kw):28s} rel err {err(e):.3e}  ({time.time()-t:.1f}s)")
