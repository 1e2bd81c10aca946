import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import graph_neural_net_b200 as pkg
import bench
G = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = bench.WORKLOADS["cfg2_er_n200_c32_b128_fwd"]
node_emb = dict(type="node_embedding", block_init="block_emb", block_inside="block", num_blocks=4, in_features=32, out_features=32, depth_of_mlp=3)
model = pkg.models.Siamese_Node_Exp(2, node_emb)
model.load_state_dict(bench.make_state_dict(cfg))
model = model.cuda().set_precision("fp16")
first = int(os.environ.get("FIRST", "0"))
x1, x2 = bench.make_inputs(cfg, first + G, seed=100)
x1, x2 = x1[first:], x2[first:]
x1 = x1.cuda()
with torch.no_grad():
    e_inf = model.node_embedder.forward_fused(x1, "fp16")
e_tr = model.node_embedder.forward_fused_train(x1, "fp16")
torch.cuda.synchronize()
print("inference finite", bool(torch.isfinite(e_inf).all()), "train finite", bool(torch.isfinite(e_tr).all()))
bad = (~torch.isfinite(e_tr)).flatten(1).any(1).nonzero().flatten().tolist()
print("graphs with non-finite training embeddings:", bad[:20], len(bad))
d = (e_tr.detach() - e_inf).flatten(1).norm(dim=1) / e_inf.flatten(1).norm(dim=1)
print("per-graph rel diff train vs inference: max", float(d[torch.isfinite(d)].max()), "argmax", int(torch.nan_to_num(d, nan=1e9).argmax()))
if len(sys.argv) > 2:
    g = torch.randn_like(e_tr)
    e_tr.backward(g)
    torch.cuda.synchronize()
    gn = sum(float(p.grad.norm()) for p in model.parameters())
    print("backward ok, sum of grad norms", gn)

if len(sys.argv) > 3:
    from graph_neural_net_b200.toolbox.losses import triplet_loss
    from graph_neural_net_b200 import _ops
    x2 = x2.cuda()
    def grads(prec):
        model.set_precision(prec)
        model.zero_grad()
        scores = model({"input": x1}, {"input": x2})
        loss = triplet_loss()(scores)
        loss.backward()
        torch.cuda.synchronize()
        return float(loss.detach()), {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    l32, g32 = grads("fp32")
    print("fp32 gradient norm", float(torch.cat([v.flatten() for v in g32.values()]).norm()))
    ref = torch.cat([v.flatten() for k, v in g32.items() if not k.endswith("convs.2.bias")])
    for lg in (9, 6, 3, 0, -3, -6):
        _ops.GradScale.log2 = lg
        l16, g16 = grads("fp16")
        nonfin = sum(int((~torch.isfinite(v)).any()) for v in g16.values())
        mine = torch.cat([v.flatten() for k, v in g16.items() if not k.endswith("convs.2.bias")])
        err = float((mine - ref).norm() / ref.norm()) if nonfin == 0 else float("nan")
        print(f"grad_scale_log2 {lg:3d}: loss {l16:.5f} (fp32 {l32:.5f}) tensors with non-finite entries {nonfin}, all-parameter rel err vs fp32 CUDA path {err:.3e}")
