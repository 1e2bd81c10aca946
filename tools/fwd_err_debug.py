import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import graph_neural_net_b200 as pkg
import bench
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = bench.WORKLOADS["cfg2_er_n200_c32_b128_fwd"]
node_emb = dict(type="node_embedding", block_init="block_emb", block_inside="block", num_blocks=4, in_features=32, out_features=32, depth_of_mlp=3)
model = pkg.models.Siamese_Node_Exp(2, node_emb)
model.load_state_dict(bench.make_state_dict(cfg))
model = model.cuda()
x1, x2 = bench.make_inputs(cfg, G, seed=100)
x1 = x1.cuda()
with torch.no_grad():
    e32 = model.node_embedder.forward_fused(x1, "fp32")
    e16 = model.node_embedder.forward_fused(x1, "fp16")
    eb = model.node_embedder.forward_fused(x1, "bf16")
d16 = (e16 - e32).flatten(1).norm(dim=1) / e32.flatten(1).norm(dim=1)
db = (eb - e32).flatten(1).norm(dim=1) / e32.flatten(1).norm(dim=1)
print("per-graph emb rel err fp16:", [f"{v:.2e}" for v in d16.tolist()])
print("per-graph emb rel err bf16:", [f"{v:.2e}" for v in db.tolist()])
# per-channel error of the worst graph
w = int(d16.argmax())
pc = (e16[w] - e32[w]).norm(dim=1) / e32[w].norm(dim=1).clamp_min(1e-12)
print("worst graph", w, "per-channel rel err:", [f"{v:.1e}" for v in pc.tolist()])
print("worst graph per-channel |e32|:", [f"{v:.1e}" for v in e32[w].norm(dim=1).tolist()])
