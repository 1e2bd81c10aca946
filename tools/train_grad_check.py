"""Per-parameter gradient errors of the 16-bit training path against a reference golden (bring-up aid).
usage: python tools/train_grad_check.py [golden name] [precision]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graph_neural_net_b200 as pkg
from graph_neural_net_b200.toolbox.losses import triplet_loss
from tests.helpers import load_golden, rel_fro
from tests.test_gpu_tc_train import build_model, feats, unpack_adj

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1_er50_c32"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
z = load_golden(name)
n = int(z["meta"][0])
model = build_model(z, prec)
if "W1" in z:
    x1, x2 = feats(z["W1"]).cuda(), feats(z["W2"]).cuda()
else:
    x1, x2 = feats(unpack_adj(z["W1_bits"], n)).cuda(), feats(unpack_adj(z["W2_bits"], n)).cuda()
scores = model({"input": x1}, {"input": x2})
loss = triplet_loss("mean")(scores)
loss.backward()
print("loss", float(loss.detach()), "ref", float(z["loss_mean"]))
ref = {k[5:]: z[k] for k in z if k.startswith("grad/")}
if len(sys.argv) > 3 and sys.argv[3] == "emul":
    from oracle import fgnn_oracle as O
    from tests.helpers import state_dict_of
    el, eg = O.emulated16_loss_and_grads(x1.cpu(), x2.cpu(), state_dict_of(z), {"fp16": torch.float16, "bf16": torch.bfloat16}[prec])
    print("emulated loss", el)
    ref = {k: v.numpy() for k, v in eg.items()}
for k, p in model.named_parameters():
    g = ref[k]
    gn = float(np.linalg.norm(g))
    mine = float(p.grad.norm())
    e = rel_fro(p.grad.cpu(), g) if gn > 0 else float("nan")
    cos = float((p.grad.cpu().flatten() @ torch.from_numpy(np.asarray(g)).flatten().float()) / max(mine * gn, 1e-30))
    print(f"{k:55s} |ref| {gn:10.3e} |ours| {mine:10.3e} rel {e:9.3e} cos {cos:+.4f}")
