"""GPU parity tests of the fp32 CUDA operators (through the C ABI / Python mirror) against the
oracle and the golden vectors produced by the reference.  Run with `pytest -m gpu` on the B200."""
import numpy as np
import pytest
import torch

import graph_neural_net_b200 as pkg
from graph_neural_net_b200.maskedtensors import maskedtensor as mt
from graph_neural_net_b200.models.layers import (MlpBlock_Real, GraphNorm, normalize, Matmul,
                                                 ColumnMaxPooling, Concat)
from graph_neural_net_b200.toolbox.losses import triplet_loss
from graph_neural_net_b200.toolbox.metrics import accuracy_max, accuracy_linear_assignment
from oracle import fgnn_oracle as O
from tests.helpers import load_golden, state_dict_of, rel_fro

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_TOL = 1e-4          # north_star: fp32 mode within 1e-4 relative on node embeddings


def feats(W):
    return torch.stack([O.adjacency_to_features(torch.from_numpy(w.astype(np.float32))) for w in W])


def build_model(z, precision="fp32"):
    n, c, nb, depth, _ = [int(v) for v in z["meta"]]
    node_emb = dict(type="node_embedding", block_init="block_emb", block_inside="block", num_blocks=nb,
                    in_features=c, out_features=c, depth_of_mlp=depth)
    model = pkg.models.Siamese_Node_Exp(2, node_emb)
    model.load_state_dict(state_dict_of(z))
    return model.to(DEV).set_precision(precision)


def test_library_loaded_and_device():
    lib = pkg.get_lib()
    assert lib.fgnn_device_supports_tcgen05() == 1


def test_layers_match_reference_fixture():
    z = load_golden("layers_f16")
    sizes = [int(s) for s in z["sizes"]]
    mlp = MlpBlock_Real(16, 32, 2)
    mlp.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("mlp/")})
    gn = GraphNorm(16)
    gn.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("gn/")})
    mlp, gn = mlp.to(DEV), gn.to(DEV)
    xs = [torch.from_numpy(z[f"x/{i}"]) for i in range(len(sizes))]
    x2s = [torch.from_numpy(z[f"x2/{i}"]) for i in range(len(sizes))]
    with torch.no_grad():
        for i, n in enumerate(sizes):
            x = xs[i][None].to(DEV)
            assert rel_fro(mlp(x)[0].cpu(), z[f"mlp_out/{i}"]) < 2e-5
            assert rel_fro(gn(x)[0].cpu(), z[f"gn_out/{i}"]) < 2e-5
            assert rel_fro(normalize(x)[0].cpu(), z[f"normalize_out/{i}"]) < 2e-5
            assert rel_fro(Matmul()(x, x2s[i][None].to(DEV))[0].cpu(), z[f"matmul_out/{i}"]) < 2e-5
            assert torch.equal(ColumnMaxPooling()(x)[0].cpu(), torch.from_numpy(z[f"colmax_out/{i}"]))
        # masked batch == per-graph results, exact zeros in the padding (reference test idiom)
        m = mt.from_list(xs, dims=(1, 2)).to(DEV)
        m2 = mt.from_list(x2s, dims=(1, 2)).to(DEV)
        mlp_m = MlpBlock_Real(16, 32, 2, constant_n_vertices=False).to(DEV)
        mlp_m.load_state_dict(mlp.state_dict())
        out = mlp_m(m)
        assert isinstance(out, mt.MaskedTensor) and out.tensor.names == ('B', None, 'N', 'N_')
        assert rel_fro(out.tensor.rename(None).cpu(), z["masked_mlp_out"]) < 2e-5
        nm = normalize(m, constant_n_vertices=False).tensor.rename(None).cpu()
        assert rel_fro(nm, z["masked_normalize_out"]) < 2e-5
        mm = Matmul()(m, m2).tensor.rename(None).cpu()
        cm = ColumnMaxPooling()(m)
        assert cm.tensor.names == ('B', None, 'N')
        for i, n in enumerate(sizes):
            assert float(out.tensor.rename(None)[i, :, n:, :].abs().sum()) == 0
            assert float(out.tensor.rename(None)[i, :, :, n:].abs().sum()) == 0
            assert rel_fro(mm[i, :, :n, :n], z[f"matmul_out/{i}"]) < 2e-5
            assert float(mm[i, :, n:, :].abs().sum()) == 0 and float(mm[i, :, :, n:].abs().sum()) == 0
            assert torch.equal(cm.tensor.rename(None)[i, :, :n].cpu(), torch.from_numpy(z[f"colmax_out/{i}"]))
            assert float(cm.tensor.rename(None)[i, :, n:].abs().sum()) == 0
        sc = mt.from_list([torch.from_numpy(z[f"score/{i}"]) for i in range(len(sizes))], dims=(0, 1)).to(DEV)
        assert abs(float(triplet_loss("mean")(sc)) - float(z["loss_mean"])) < 1e-5
        assert abs(float(triplet_loss("mean_of_mean")(sc)) - float(z["loss_mean_of_mean"])) < 1e-5
        assert list(accuracy_max(sc)) == list(z["acc"])


@pytest.mark.parametrize("name", ["tiny_er12_c8", "cfg1_er50_c32", "cfg3_reg40_c64"])
def test_fp32_model_matches_reference(name):
    z = load_golden(name)
    model = build_model(z)
    x1, x2 = feats(z["W1"]).to(DEV), feats(z["W2"]).to(DEV)
    with torch.no_grad():
        outs = model.node_embedder({"input": x1})
        e1 = outs["ne/suffix"]
        e2 = model.embed({"input": x2})
        scores = model({"input": x1}, {"input": x2})
        assert rel_fro(e1.cpu(), z["emb1"]) < FP32_TOL and rel_fro(e2.cpu(), z["emb2"]) < FP32_TOL
        assert rel_fro(scores.cpu(), z["scores"]) < FP32_TOL
        assert abs(float(triplet_loss("mean")(scores)) - float(z["loss_mean"])) < 1e-4
        assert abs(float(triplet_loss("mean_of_mean")(scores)) - float(z["loss_mean_of_mean"])) < 1e-4
        # per-stage parity: every intermediate node the reference's Network.forward returns
        for k, v in z.items():
            if k.startswith("tap/"):
                assert rel_fro(outs[k[4:]][0, :4].cpu(), v) < FP32_TOL, k
        # fused fp32 entry point == node-by-node execution
        fused = model.node_embedder.forward_fused(x1, "fp32")
        assert rel_fro(fused.cpu(), e1.cpu()) < 1e-6
    # argmax / accuracy parity on the reference's own scores (bit-identical inputs => identical counts)
    ref_scores = torch.from_numpy(z["scores"]).to(DEV)
    assert list(accuracy_max(ref_scores)) == list(z["acc"])
    # the reference's default metric (trainers.py:53,74): device LAP == the reference's scipy result on these scores
    lap_ref = load_golden("lap_acc")
    assert list(accuracy_linear_assignment(ref_scores)) == list(lap_ref[name + "/acc_lap"])
    each = accuracy_linear_assignment(ref_scores, aggregate_score=False)
    assert np.allclose(each, lap_ref[name + "/acc_lap_each"])


def test_ragged_batch_matches_per_graph_reference():
    z = load_golden("ragged_c16")
    sizes = [int(s) for s in z["sizes"]]
    nmax, c, nb, depth, _ = [int(v) for v in z["meta"]]
    node_emb = dict(type="node_embedding", block_init="block_emb", block_inside="block", num_blocks=nb,
                    in_features=c, out_features=c, depth_of_mlp=depth, constant_n_vertices=False)
    model = pkg.models.Siamese_Node_Exp(2, node_emb)
    model.load_state_dict(state_dict_of(z))
    model = model.to(DEV)
    g1 = [O.adjacency_to_features(torch.from_numpy(z[f"W1/{i}"].astype(np.float32))) for i in range(len(sizes))]
    g2 = [O.adjacency_to_features(torch.from_numpy(z[f"W2/{i}"].astype(np.float32))) for i in range(len(sizes))]
    m1 = mt.from_list(g1, dims=(1, 2), base_name='N').to(DEV)
    m2 = mt.from_list(g2, dims=(1, 2), base_name='N').to(DEV)
    with torch.no_grad():
        e1 = model.embed({"input": m1})
        assert rel_fro(e1.tensor.rename(None).cpu(), z["masked_emb1"]) < FP32_TOL
        scores = model({"input": m1}, {"input": m2})
        assert isinstance(scores, mt.MaskedTensor)
        for i, n in enumerate(sizes):
            assert rel_fro(scores[i].cpu(), z[f"scores/{i}"]) < FP32_TOL
            assert float(e1.tensor.rename(None)[i, :, n:].abs().sum()) == 0
        assert abs(float(triplet_loss("mean")(scores)) - float(z["loss_mean"])) < 1e-4
        assert abs(float(triplet_loss("mean_of_mean")(scores)) - float(z["loss_mean_of_mean"])) < 1e-4
        fused = model.node_embedder.forward_fused(m1, "fp32")
        assert rel_fro(fused.tensor.rename(None).cpu(), z["masked_emb1"]) < FP32_TOL


def test_head_backward_matches_autograd():
    """loss -> scores -> embeddings gradients of the fused head vs torch autograd on the oracle."""
    gen = torch.Generator().manual_seed(5)
    e1 = torch.randn((3, 8, 21), generator=gen)
    e2 = torch.randn((3, 8, 21), generator=gen)
    a1, a2 = e1.clone().requires_grad_(True), e2.clone().requires_grad_(True)
    O.triplet_loss(O.siamese_scores(a1, a2)).backward()
    b1, b2 = e1.to(DEV).requires_grad_(True), e2.to(DEV).requires_grad_(True)
    from graph_neural_net_b200 import _ops
    loss = triplet_loss()(_ops.ScoresFunction.apply(b1, b2, None))
    loss.backward()
    assert rel_fro(b1.grad.cpu(), a1.grad) < 1e-5 and rel_fro(b2.grad.cpu(), a2.grad) < 1e-5


def test_cpu_tensors_are_rejected():
    with pytest.raises(pkg.FgnnError):
        Matmul()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))


@pytest.mark.parametrize("name", ["tiny_er12_c8", "cfg1_er50_c32"])
def test_fp32_training_gradients_match_reference_autograd(name):
    """loss.backward() through every fp32 CUDA operator vs the gradients the reference's autograd
    produced for the same weights and inputs (tests/golden, written by oracle/make_golden.py)."""
    z = load_golden(name)
    model = build_model(z)
    x1, x2 = feats(z["W1"]).to(DEV), feats(z["W2"]).to(DEV)
    scores = model({"input": x1}, {"input": x2})
    loss = triplet_loss("mean")(scores)
    assert abs(float(loss) - float(z["loss_mean"])) < 1e-4
    loss.backward()
    worst = 0.0
    for k, p in model.named_parameters():
        g = z["grad/" + k]
        assert p.grad is not None, k
        if np.abs(g).max() < 1e-6:                       # last-conv biases: mathematically zero
            assert float(p.grad.abs().max()) < 1e-5, k
        else:
            e = rel_fro(p.grad.cpu(), g)
            worst = max(worst, e)
            assert e < 5e-3, (k, e)
    print(f"{name}: worst parameter-gradient rel err {worst:.2e}")


def test_fp32_train_step_reduces_loss():
    """A few Adam steps of graph_neural_net_b200.training.train_step (world size 1) on one batch."""
    from graph_neural_net_b200.training import train_step
    z = load_golden("tiny_er12_c8")
    model = build_model(z)
    opt = model.configure_optimizers()["optimizer"]
    x1, x2 = feats(z["W1"]).to(DEV), feats(z["W2"]).to(DEV)
    losses = [train_step(model, opt, {"input": x1}, {"input": x2})[0] for _ in range(8)]
    assert abs(losses[0] - float(z["loss_mean"])) < 1e-4
    assert losses[-1] < losses[0]


def test_features_from_adjacency_matches_reference_representation():
    """input construction on the device == the reference's host construction + padding."""
    from graph_neural_net_b200.loaders import data_generator as dg
    gen = torch.Generator().manual_seed(5)
    sizes = [37, 64, 5, 50]
    N = max(sizes)
    adj = torch.zeros((len(sizes), N, N), dtype=torch.uint8)
    graphs = []
    for g, n in enumerate(sizes):
        W = (torch.rand((n, n), generator=gen) < 0.3)
        W = torch.triu(W, 1)
        W = (W | W.T)
        adj[g, :n, :n] = W.to(torch.uint8)
        graphs.append(O.adjacency_to_features(W.float()))
    # garbage outside the n x n blocks must be ignored
    adj[0, 40:, :] = 1
    ref = mt.from_list(graphs, dims=(1, 2)).tensor.rename(None)
    out = dg.adjacency_batch_to_tensor_representation(adj.to(DEV), torch.tensor(sizes, dtype=torch.int32, device=DEV))
    assert torch.equal(out.cpu(), ref)
    dense = dg.adjacency_batch_to_tensor_representation(adj[1:2, :64, :64].contiguous().to(DEV))
    assert torch.equal(dense[0].cpu(), graphs[1])


@pytest.mark.parametrize("n,sizes", [(12, None), (50, None), (200, None), (64, [64, 1, 33, 2]), (500, [500, 257])])
def test_device_linear_assignment_equals_scipy(n, sizes):
    """fgnn_lap_fwd: identical assignment and optimal cost to scipy.optimize.linear_sum_assignment
    on -log_softmax(scores), the reference's call (toolbox/metrics.py:100-106); padded rows report -1."""
    from scipy.optimize import linear_sum_assignment
    from graph_neural_net_b200.toolbox.metrics import linear_assignment
    gen = torch.Generator().manual_seed(n)
    G = len(sizes) if sizes else 3
    scores = torch.randn((G, n, n), generator=gen) * 3
    scores[0] += 4 * torch.eye(n)                      # one graph with a strong diagonal (trained-like margins)
    x = scores.to(DEV)
    if sizes:
        x = mt.from_list([scores[i, :s, :s] for i, s in enumerate(sizes)], dims=(0, 1)).to(DEV)
    cols, correct, cost = linear_assignment(x)
    cols, correct, cost = cols.cpu().numpy(), correct.cpu().numpy(), cost.cpu().numpy()
    for g in range(G):
        m = sizes[g] if sizes else n
        ref_cost_mat = -torch.log_softmax(scores[g, :m, :m], -1).numpy().astype(np.float64)
        rows, preds = linear_sum_assignment(ref_cost_mat)
        assert np.array_equal(cols[g, :m], preds), g
        assert np.all(cols[g, m:] == -1)
        assert correct[g] == int(np.sum(preds == np.arange(m)))
        assert abs(cost[g] - ref_cost_mat[rows, preds].sum()) < 1e-5 * max(1.0, abs(cost[g]))


def test_lap_on_trained_model_scores_matches_reference():
    z = load_golden("trained_er50_c32")
    from graph_neural_net_b200.toolbox.metrics import linear_assignment
    scores = torch.from_numpy(z["scores"]).to(DEV)
    cols, correct, _ = linear_assignment(scores)
    assert np.array_equal(cols.cpu().numpy(), z["lap_preds"])
    assert list(accuracy_linear_assignment(scores)) == list(z["acc_lap"])
    assert list(accuracy_max(scores)) == list(z["acc"])


def test_lightning_step_functions_run_the_reference_recipe():
    """Siamese_Node_Exp.training_step / validation_step / test_step / configure_optimizers (reference models/trainers.py:70-104):
    the step computes triplet_loss and the LAP accuracy on the model's scores and logs both; training_step's loss is the
    reference's loss on the golden batch and it back-propagates; the optimiser recipe is Adam + ReduceLROnPlateau on val_loss."""
    z = load_golden("cfg1_er50_c32")
    model = build_model(z)
    logged = {}
    model.log = lambda k, v, *a, **kw: logged.__setitem__(k, float(v))
    x1, x2 = feats(z["W1"]).to(DEV), feats(z["W2"]).to(DEV)
    batch = ({"input": x1}, {"input": x2})
    loss = model.training_step(batch, 0)
    assert abs(float(loss) - float(z["loss_mean"])) < 1e-4
    lap_ref = load_golden("lap_acc")
    acc, tot = [int(v) for v in lap_ref["cfg1_er50_c32/acc_lap"]]
    assert abs(logged["train_loss"] - float(z["loss_mean"])) < 1e-4 and abs(logged["train_acc"] - acc / tot) < 1e-6
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    with torch.no_grad():
        assert model.validation_step(batch, 0) is None and model.test_step(batch, 0) is None
    assert abs(logged["val_loss"] - logged["train_loss"]) < 1e-6 and abs(logged["test_acc"] - acc / tot) < 1e-6
    opt = model.configure_optimizers()
    assert isinstance(opt["optimizer"], torch.optim.Adam) and opt["lr_scheduler"]["monitor"] == "val_loss"
    sch = opt["lr_scheduler"]["scheduler"]
    assert isinstance(sch, torch.optim.lr_scheduler.ReduceLROnPlateau) and sch.factor == model.scheduler_decay \
        and sch.patience == model.scheduler_step and sch.min_lrs == [model.lr_stop]


def test_device_generators_follow_the_reference_distributions():
    """fgnn_generate_pairs_u8 against the definitions of the reference generators
    (loaders/data_generator.py:39-87).  Parity is distributional: simple symmetric graphs, the ER edge density, EXACT
    degrees for "Regular" (d = int(p n), +1 if n d odd), the two flip rates of noise_erdos_renyi, and a triangle count that
    matches a random (not a circulant) regular graph; the same seed reproduces the batch."""
    from graph_neural_net_b200.loaders.data_generator import generate_pairs_on_device
    p, noise = 0.2, 0.1
    for gen, n in (("ErdosRenyi", 200), ("Regular", 200), ("Regular", 51)):
        G = 6
        a1, a2 = generate_pairs_on_device(gen, G, n, p, noise, seed=3787)
        b1, b2 = generate_pairs_on_device(gen, G, n, p, noise, seed=3787)
        c1, _ = generate_pairs_on_device(gen, G, n, p, noise, seed=3788)
        assert torch.equal(a1, b1) and torch.equal(a2, b2) and not torch.equal(a1, c1)
        for a in (a1, a2):
            assert a.dtype == torch.uint8 and int(a.max()) == 1
            assert torch.equal(a, a.transpose(1, 2)) and int(a.diagonal(dim1=1, dim2=2).sum()) == 0
        w = a1.double()
        pairs_tot = G * n * (n - 1)
        if gen == "ErdosRenyi":
            dens = float(w.sum()) / pairs_tot
            assert abs(dens - p) < 4 * (p * (1 - p) / (pairs_tot / 2)) ** 0.5, dens
            d_eff = p
        else:
            d = int(p * n)
            d += (n * d) % 2
            assert torch.equal(a1.sum(2), torch.full((G, n), d, dtype=a1.sum(2).dtype, device=a1.device))      # exactly d-regular
            assert not torch.equal(a1[0], a1[1])
            d_eff = d / (n - 1)
            tri = float(torch.einsum("gij,gjk,gki->", w, w, w)) / 6 / G
            expect = n * (n - 1) * (n - 2) / 6 * d_eff * ((d - 1) / (n - 2)) ** 2     # first-order count for a random d-regular graph
            assert abs(tri - expect) < 0.15 * expect, (tri, expect)         # the circulant start has more than twice as many
        removed = float(((a1 == 1) & (a2 == 0)).sum()) / float(w.sum())
        added = float(((a1 == 0) & (a2 == 1)).sum()) / (pairs_tot - float(w.sum()))
        assert abs(removed - noise) < 0.01 and abs(added - d_eff * noise / (1 - d_eff)) < 0.004, (removed, added)
    # ragged sizes inside the padding, fed to the embedder's adjacency entry point
    sizes = torch.tensor([30, 64, 17], dtype=torch.int32, device=DEV)
    a1, a2 = generate_pairs_on_device("ErdosRenyi", 3, 64, p, noise, seed=5, sizes=sizes)
    for g, ng in enumerate([30, 64, 17]):
        assert int(a1[g, ng:, :].sum()) == 0 and int(a1[g, :, ng:].sum()) == 0 and int(a2[g, ng:, :].sum()) == 0
    with pytest.raises(NotImplementedError):
        generate_pairs_on_device("BarabasiAlbert", 1, 10, p, noise)


def test_ragged_gradients_match_reference_per_graph_autograd():
    """cfg4's backward: gradients of the ragged loss (sum_b CE_b / sum_b n_b, toolbox/losses.py:27-34) through the masked fp32
    CUDA operators on a MaskedTensor batch vs the reference's autograd over its per-graph dense loop (fixture
    ragged_cstn_c16 `grad/*`, written by oracle/make_golden.py from the unmodified reference)."""
    from graph_neural_net_b200.maskedtensors import maskedtensor as mt
    from graph_neural_net_b200.toolbox.losses import triplet_loss
    from oracle import fgnn_oracle as O
    z = load_golden("ragged_cstn_c16")
    nmax, c, nb, depth, _ = [int(v) for v in z["meta"]]
    sizes = [int(v) for v in z["sizes"]]
    node_emb = dict(type="node_embedding", block_init="block_emb", block_inside="block", num_blocks=nb,
                    in_features=c, out_features=c, depth_of_mlp=depth, constant_n_vertices=False)
    model = pkg.models.Siamese_Node_Exp(2, node_emb)
    model.load_state_dict(state_dict_of(z))
    model = model.to(DEV).set_precision("fp32")
    g1 = [O.adjacency_to_features(torch.from_numpy(z[f"W1/{i}"].astype(np.float32))) for i in range(len(sizes))]
    g2 = [O.adjacency_to_features(torch.from_numpy(z[f"W2/{i}"].astype(np.float32))) for i in range(len(sizes))]
    x1 = mt.from_list(g1, dims=(1, 2)).to(DEV)
    x2 = mt.from_list(g2, dims=(1, 2)).to(DEV)
    loss = triplet_loss("mean")(model({"input": x1}, {"input": x2}))
    assert abs(float(loss.detach()) - float(z["loss_mean"])) < 1e-4
    loss.backward()
    worst = 0.0
    for k, p in model.named_parameters():
        g = z["grad/" + k]
        if k.endswith(f"convs.{depth - 1}.bias"):
            assert float(p.grad.abs().max()) < 1e-5, k          # cancels in GraphNorm
            continue
        worst = max(worst, rel_fro(p.grad.cpu(), g))
    print(f"PARITY ragged fp32 gradients vs reference per-graph autograd: worst per-tensor rel err {worst:.2e}")
    assert worst < 2e-4
