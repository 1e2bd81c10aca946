"""GPU bring-up checks for the tensor-core kernels, one stage per process (a device trap poisons
the CUDA context).  Usage: python tools/tc_check.py <stage> [precision]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import graph_neural_net_b200 as pkg
from graph_neural_net_b200 import _lib as L, _ops
from oracle import fgnn_oracle as O

dev = "cuda:0"
lib = pkg.get_lib()
stage = sys.argv[1]
prec_name = sys.argv[2] if len(sys.argv) > 2 else "bf16"
prec = L.PRECISIONS[prec_name]
tdt = torch.bfloat16 if prec_name == "bf16" else torch.float16


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def tc_matmul(a, b, n_dev=None):
    G, Cc, N, _ = a.shape
    out = torch.empty_like(a)
    nb = lib.fgnn_debug_tc_matmul_workspace_bytes(G, Cc, N)
    ws = L.workspace(a.device, nb)
    L.check(lib.fgnn_debug_tc_matmul(prec, L.ptr(a), L.ptr(b), L.ptr(out), G, Cc, N,
                                     L.ptr(n_dev) if n_dev is not None else None, L.ptr(ws), ws.numel(),
                                     L.stream_ptr(a.device)), "debug_tc_matmul")
    torch.cuda.synchronize()
    return out


def check_matmul(G, Cc, N, sizes=None):
    gen = torch.Generator().manual_seed(N)
    a = torch.randn((G, Cc, N, N), generator=gen).to(dev)
    b = torch.randn((G, Cc, N, N), generator=gen).to(dev)
    n_dev = None
    if sizes is not None:
        n_dev = torch.tensor(sizes, dtype=torch.int32, device=dev)
    out = tc_matmul(a, b, n_dev)
    ar, br = a.to(tdt).float(), b.to(tdt).float()
    worst = 0.0
    for g in range(G):
        n = N if sizes is None else sizes[g]
        ref = torch.matmul(ar[g, :, :n, :n].double(), br[g, :, :n, :n].double())
        e = rel(out[g, :, :n, :n], ref)
        worst = max(worst, e)
        if sizes is not None:
            assert float(out[g, :, n:, :].abs().sum()) == 0 and float(out[g, :, :, n:].abs().sum()) == 0
    print(f"matmul[{prec_name}] G={G} C={Cc} N={N} sizes={sizes}: worst rel err {worst:.3e}")
    assert worst < 6e-3, worst


def check_mlp(c_in, c_out, depth, G, N, sizes=None):
    gen = torch.Generator().manual_seed(7 * N + c_out)
    x = torch.randn((G, c_in, N, N), generator=gen)
    ws_ = [torch.randn((c_out, c_in if k == 0 else c_out), generator=gen) / (c_in if k == 0 else c_out) ** 0.5
           for k in range(depth)]
    bs = [torch.randn(c_out, generator=gen) * 0.1 for _ in range(depth)]
    gw = 1 + 0.3 * torch.randn(c_out, generator=gen)
    gb = 0.2 * torch.randn(c_out, generator=gen)
    sd = {}
    for k in range(depth):
        sd[f"m.convs.{k}.weight"] = ws_[k].reshape(c_out, -1, 1, 1)
        sd[f"m.convs.{k}.bias"] = bs[k]
    sd["m.gn.weight"], sd["m.gn.bias"] = gw, gb
    keep = []
    p = _ops.make_mlp_params([w.to(dev) for w in ws_], [b.to(dev) for b in bs], gw.to(dev), gb.to(dev), 1e-5, keep)
    xd = x.to(dev)
    n_dev = torch.tensor(sizes, dtype=torch.int32, device=dev) if sizes is not None else None
    if sizes is not None:
        for g, n in enumerate(sizes):
            xd[g, :, n:, :] = 0
            xd[g, :, :, n:] = 0
    y = torch.empty((G, c_out, N, N), device=dev)
    nb = lib.fgnn_debug_tc_mlp_workspace_bytes(G, c_in, c_out, depth, N)
    wsb = L.workspace(dev, nb)
    L.check(lib.fgnn_debug_tc_mlp(prec, C.byref(p), L.ptr(xd), L.ptr(y), G, N,
                                  L.ptr(n_dev) if n_dev is not None else None, L.ptr(wsb), wsb.numel(),
                                  L.stream_ptr(dev)), "debug_tc_mlp")
    torch.cuda.synchronize()
    worst = 0.0
    for g in range(G):
        n = N if sizes is None else sizes[g]
        ref = O.mlp_block(x[g:g + 1, :, :n, :n].double(), {k: v.double() for k, v in sd.items()}, "m", depth)[0]
        worst = max(worst, rel(y[g, :, :n, :n].cpu(), ref))
    print(f"mlp[{prec_name}] {c_in}->{c_out} depth={depth} G={G} N={N} sizes={sizes}: worst rel err {worst:.3e}")
    assert worst < (3e-2 if prec_name == "bf16" else 4e-3), worst


def check_embed(n, c, pairs, sizes=None, reg=False):
    gen = torch.Generator().manual_seed(3787)
    sd = O.xavier_state_dict(2, c, 4, 3, gen, randomize_gn=True)
    node_emb = dict(type="node_embedding", block_init="block_emb", block_inside="block", num_blocks=4,
                    in_features=c, out_features=c, depth_of_mlp=3)
    model = pkg.models.Siamese_Node_Exp(2, node_emb).to(dev)
    model.load_state_dict(sd)
    if sizes is None:
        xs = torch.stack([O.synthetic_pair(n, 0.2, 0.1, gen, regular_degree=int(0.2 * n) if reg else None)[0]
                          for _ in range(pairs)])
        ref = O.node_embedding(xs.double(), {k: v.double() for k, v in sd.items()})
        inp = xs.to(dev)
    else:
        from graph_neural_net_b200.maskedtensors import maskedtensor as mt
        graphs = [O.synthetic_pair(s, 0.3, 0.1, gen)[0] for s in sizes]
        refs = O.node_embedding_ragged([g.double() for g in graphs], {k: v.double() for k, v in sd.items()})
        inp = mt.from_list(graphs, dims=(1, 2)).to(dev)
    with torch.no_grad():
        e32 = model.node_embedder.forward_fused(inp, "fp32")
        e = model.node_embedder.forward_fused(inp, prec_name)
    torch.cuda.synchronize()
    if sizes is None:
        print(f"embed[{prec_name}] n={n} c={c}: rel err vs fp64 oracle {rel(e.cpu(), ref):.3e} "
              f"(fp32 path {rel(e32.cpu(), ref):.3e})")
    else:
        et = e.tensor.rename(None).cpu()
        worst = max(rel(et[i, :, :s], refs[i]) for i, s in enumerate(sizes))
        pad = max(float(et[i, :, s:].abs().sum()) for i, s in enumerate(sizes))
        print(f"embed[{prec_name}] ragged {sizes} c={c}: worst rel err {worst:.3e}, padding abs sum {pad}; per graph",
              [f"{rel(et[i, :, :s], refs[i]):.2e}" for i, s in enumerate(sizes)],
              "finite", [bool(torch.isfinite(et[i]).all()) for i in range(len(sizes))])


if stage == "matmul64":
    check_matmul(2, 3, 40)
    check_matmul(1, 2, 64)
elif stage == "matmul128":
    check_matmul(2, 2, 100)
elif stage == "matmul256":
    check_matmul(1, 2, 200)
    check_matmul(1, 1, 500)
elif stage == "matmul_ragged":
    check_matmul(3, 2, 150, sizes=[150, 70, 33])
elif stage == "mlp64":
    check_mlp(64, 64, 3, 2, 40)
    check_mlp(2, 64, 3, 1, 50)
    check_mlp(128, 64, 3, 1, 72)
elif stage == "mlp32":
    check_mlp(32, 32, 3, 2, 40)
    check_mlp(2, 32, 2, 2, 30, sizes=[30, 17])
    check_mlp(34, 32, 3, 1, 50)
elif stage == "embed":
    check_embed(40, 32, 3)
    check_embed(50, 64, 2)
    check_embed(100, 64, 2, reg=True)
elif stage == "embed_ragged":
    check_embed(0, 32, 0, sizes=[50, 23, 37, 64])
    check_embed(0, 64, 0, sizes=[130, 70])
elif stage == "ragged_probe":
    for sizes in ([130], [130, 130], [70, 130], [130, 70], [100, 70], [200, 150], [129, 128], [140, 127]):
        try:
            check_embed(0, 64, 0, sizes=sizes)
        except Exception as ex:
            print("FAILED", sizes, ex)
else:
    raise SystemExit("unknown stage")
print("STAGE", stage, prec_name, "OK")
