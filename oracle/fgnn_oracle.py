"""CPU oracle for the 2-FGNN siamese hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain functional restatement (torch tensor algebra, no nn.Module, no reference
imports) of what mlelarge/graph_neural_net computes on the path named by
BASELINE.json.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import this module.  The product path
(graph_neural_net_b200/*) never does and raises if its CUDA library is missing.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function below
against fixtures in tests/golden/*.npz that were produced by running the unmodified
reference (imported through oracle/refshim.py) on seeded inputs; the generating
script is oracle/make_golden.py.

Each function cites the reference lines it restates (paths relative to
/root/reference).  All maths runs in the dtype of the inputs (fp32 by default,
fp64 if the caller up-casts) on whatever device the inputs live on.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------- #
# GraphNorm / normalize                         models/layers.py:47-80
# --------------------------------------------------------------------------- #
def normalize(y: Tensor, n: Optional[float] = None, eps: float = 1e-5) -> Tensor:
    """(y - mean) / (2*sqrt(n*(var+eps))), mean/var (biased) over the last two dims.

    models/layers.py:71-80.  `n` is the vertex count (y.size(-1) when the batch has a
    constant number of vertices, the per-graph size in the masked case).
    """
    if n is None:
        n = y.size(-1)
    mu = y.mean(dim=(-1, -2), keepdim=True)
    var = ((y - mu) ** 2).mean(dim=(-1, -2), keepdim=True)      # unbiased=False
    return (y - mu) / (2.0 * torch.sqrt(n * (var + eps)))


def graph_norm(y: Tensor, weight: Tensor, bias: Tensor, n: Optional[float] = None,
               eps: float = 1e-5) -> Tensor:
    """weight * normalize(y) + bias with (1,C,1,1) affine.  models/layers.py:68-69."""
    return weight.reshape(1, -1, 1, 1) * normalize(y, n, eps) + bias.reshape(1, -1, 1, 1)


# --------------------------------------------------------------------------- #
# MlpBlock_Real                                 models/layers.py:109-131
# --------------------------------------------------------------------------- #
def conv1x1(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """nn.Conv2d(kernel_size=1): per-pixel channel mixing.  models/layers.py:120."""
    w2 = w.reshape(w.shape[0], w.shape[1])
    y = torch.einsum("oc,bcij->boij", w2, x)
    if b is not None:
        y = y + b.reshape(1, -1, 1, 1)
    return y


def mlp_block(x: Tensor, sd: StateDict, prefix: str, depth: int,
              n: Optional[float] = None, eps: float = 1e-5) -> Tensor:
    """conv -> relu -> ... -> conv -> GraphNorm.  models/layers.py:126-131."""
    h = x
    for k in range(depth):
        h = conv1x1(h, sd[f"{prefix}.convs.{k}.weight"], sd[f"{prefix}.convs.{k}.bias"])
        if k < depth - 1:
            h = torch.relu(h)
    return graph_norm(h, sd[f"{prefix}.gn.weight"], sd[f"{prefix}.gn.bias"], n, eps)


# --------------------------------------------------------------------------- #
# block / base_model / node_embedding           models/blocks_emb.py:16-43
# --------------------------------------------------------------------------- #
def fgnn_block(x: Tensor, sd: StateDict, prefix: str, depth: int,
               n: Optional[float] = None, taps: Optional[dict] = None) -> Tensor:
    """mlp3(cat[mlp1(x) @ mlp2(x), x]).  models/blocks_emb.py:16-27,
    Matmul models/layers.py:161-162, Concat models/layers.py:145-146."""
    y1 = mlp_block(x, sd, f"{prefix}_mlp1", depth, n)
    y2 = mlp_block(x, sd, f"{prefix}_mlp2", depth, n)
    mult = torch.matmul(y1, y2)
    cat = torch.cat((mult, x), dim=1)
    out = mlp_block(cat, sd, f"{prefix}_mlp3", depth, n)
    if taps is not None:
        taps[prefix + "/mlp1"] = y1
        taps[prefix + "/mlp2"] = y2
        taps[prefix + "/mult"] = mult
        taps[prefix + "/mlp3"] = out
    return out


def count_blocks(sd: StateDict, root: str = "node_embedder.ne_bm_block") -> Tuple[int, int]:
    """(num_blocks, depth_of_mlp) recovered from state-dict keys (models/utils.py:57-58 naming)."""
    nb = 0
    while f"{root}{nb + 1}_mlp1.convs.0.weight" in sd:
        nb += 1
    depth = 0
    while f"{root}1_mlp1.convs.{depth}.weight" in sd:
        depth += 1
    return nb, depth


def node_embedding(x: Tensor, sd: StateDict, n: Optional[float] = None,
                   root: str = "node_embedder.ne_bm_block", taps: Optional[dict] = None) -> Tensor:
    """num_blocks blocks then max over the last dim -> (B,C,N).
    models/blocks_emb.py:29-43, ColumnMaxPooling models/layers.py:194-203."""
    nb, depth = count_blocks(sd, root)
    h = x
    for i in range(1, nb + 1):
        h = fgnn_block(h, sd, f"{root}{i}", depth, n, taps)
    return h.max(dim=-1)[0]


def node_embedding_ragged(graphs: Sequence[Tensor], sd: StateDict,
                          root: str = "node_embedder.ne_bm_block") -> List[Tensor]:
    """Ragged oracle = per-graph dense loop, the reference's own test idiom
    (maskedtensors/test_maskedtensor.py:22-34, :167-188): a masked batch must equal each
    un-padded graph run alone.  Returns a list of (C, n_b) embeddings."""
    return [node_embedding(g.unsqueeze(0), sd, None, root)[0] for g in graphs]


# --------------------------------------------------------------------------- #
# Siamese head                                  models/trainers.py:60-68
# --------------------------------------------------------------------------- #
def siamese_scores(e1: Tensor, e2: Tensor) -> Tensor:
    """scores[b] = e1[b]^T e2[b]: (B,C,N),(B,C,N) -> (B,N,N).  models/trainers.py:67."""
    return torch.matmul(e1.transpose(1, 2), e2)


def siamese_forward(x1: Tensor, x2: Tensor, sd: StateDict) -> Tensor:
    return siamese_scores(node_embedding(x1, sd), node_embedding(x2, sd))


# --------------------------------------------------------------------------- #
# triplet_loss                                  toolbox/losses.py:8-34
# --------------------------------------------------------------------------- #
def ce_sum_identity(scores: Tensor) -> Tensor:
    """sum_i [logsumexp_j s[i,j] - s[i,i]] for one (n,n) score matrix.
    nn.CrossEntropyLoss(reduction='sum')(out, arange(n)), toolbox/losses.py:27-31."""
    return (torch.logsumexp(scores, dim=-1) - torch.diagonal(scores, dim1=-2, dim2=-1)).sum()


def triplet_loss(score_list: Sequence[Tensor], loss_reduction: str = "mean") -> Tensor:
    """'mean': sum CE / sum n ; 'mean_of_mean': mean_b(CE_b / n_b).  toolbox/losses.py:12-34.
    `score_list` is an iterable of (n_b, n_b) matrices (a (B,N,N) tensor works too)."""
    if loss_reduction not in ("mean", "mean_of_mean"):
        raise ValueError("Unknown loss_reduction parameters {}".format(loss_reduction))
    loss = 0.0
    total = 0
    for s in score_list:
        nv = s.shape[0]
        ce = ce_sum_identity(s)
        if loss_reduction == "mean":
            loss, total = loss + ce, total + nv
        else:
            loss, total = loss + ce / nv, total + 1
    return loss / total


# --------------------------------------------------------------------------- #
# accuracy_max                                  toolbox/metrics.py:118-141
# --------------------------------------------------------------------------- #
def accuracy_max(score_list: Sequence[Tensor]) -> Tuple[int, int]:
    """(#rows whose argmax_j is the diagonal, #rows).  toolbox/metrics.py:118-141
    (labels=None, aggregate_score=True)."""
    correct = 0
    total = 0
    for s in score_list:
        pred = torch.argmax(s, dim=1)
        correct += int((pred == torch.arange(s.shape[0], device=s.device)).sum())
        total += s.shape[0]
    return correct, total


# --------------------------------------------------------------------------- #
# accuracy_linear_assignment                    toolbox/metrics.py:92-116
# --------------------------------------------------------------------------- #
def linear_assignment_preds(scores: Tensor):
    """Hungarian matching maximising sum_i log_softmax(scores)[i, pred_i] (scipy, as the reference:
    toolbox/metrics.py:100-106).  Returns the column assigned to each row."""
    from scipy.optimize import linear_sum_assignment
    cost = -torch.log_softmax(scores, -1).detach().cpu().numpy()
    return linear_sum_assignment(cost)[1]


def accuracy_linear_assignment(score_list: Sequence[Tensor]) -> Tuple[int, int]:
    """(#rows matched to their own index by the optimal assignment, #rows).  toolbox/metrics.py:92-116
    (labels=None, aggregate_score=True)."""
    import numpy as np
    correct = 0
    total = 0
    for s in score_list:
        preds = linear_assignment_preds(s)
        correct += int(np.sum(preds == np.arange(s.shape[0])))
        total += s.shape[0]
    return correct, total


# --------------------------------------------------------------------------- #
# Input construction                            loaders/data_generator.py:118-125
# --------------------------------------------------------------------------- #
def adjacency_to_features(W: Tensor) -> Tensor:
    """B[0]=W, B[1]=diag(deg).  loaders/data_generator.py:118-125."""
    n = W.shape[0]
    B = torch.zeros((2, n, n), dtype=W.dtype, device=W.device)
    B[0] = W
    B[1] = torch.diag(W.sum(1))
    return B


def pad_batch(graphs: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """Zero-pad a list of (F,n_b,n_b) tensors to (B,F,Nmax,Nmax) + int sizes.
    maskedtensors/maskedtensor.py:8-48 (prefix masks <=> one size per graph)."""
    nmax = max(int(g.shape[-1]) for g in graphs)
    out = torch.zeros((len(graphs), graphs[0].shape[0], nmax, nmax), dtype=graphs[0].dtype)
    for i, g in enumerate(graphs):
        n = g.shape[-1]
        out[i, :, :n, :n] = g
    return out, torch.tensor([int(g.shape[-1]) for g in graphs], dtype=torch.int32)


# --------------------------------------------------------------------------- #
# Algorithmic FLOPs -- used by bench.py's roofline
# --------------------------------------------------------------------------- #
def flops_per_graph(n: int, c: int, num_blocks: int = 4, depth: int = 3, c_in0: int = 2,
                    c_out: Optional[int] = None) -> Tuple[float, float]:
    """(conv FLOPs, matmul FLOPs) for one graph's embedder forward; 2 FLOP per MAC, GEMMs only."""
    c_out = c if c_out is None else c_out
    f_conv = 0.0
    f_mm = 0.0
    ci = c_in0
    for k in range(num_blocks):
        co = c if k < num_blocks - 1 else c_out
        mlp12 = 2 * (ci * co + (depth - 1) * co * co)
        mlp3 = (ci + co) * co + (depth - 1) * co * co
        f_conv += 2.0 * n * n * (mlp12 + mlp3)
        f_mm += 2.0 * co * float(n) ** 3
        ci = co
    return f_conv, f_mm


def flops_per_pair(n: int, c: int, num_blocks: int = 4, depth: int = 3, c_in0: int = 2) -> float:
    fc, fm = flops_per_graph(n, c, num_blocks, depth, c_in0)
    return 2.0 * (fc + fm) + 2.0 * n * n * c


# --------------------------------------------------------------------------- #
# Seeded synthetic inputs that need no networkx (bench / GPU-box tests)
# --------------------------------------------------------------------------- #
def synthetic_pair(n: int, p: float, noise: float, gen: torch.Generator,
                   regular_degree: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """One (clean, noisy) pair with the reference's input format.

    Distributionally equivalent to loaders/data_generator.py:39-44 (ER) and :79-87 (ER noise
    W*(1-N1)+(1-W)*N2 with N1~ER(noise), N2~ER(p*noise/(1-p))), but drawn from a torch
    Generator so it can run on the GPU box where networkx seeding is not reproducible.
    `regular_degree` builds a circulant d-regular graph with randomly relabelled vertices
    (a stand-in for random_regular_graph, :58-68: same degree sequence and density)."""
    def er(prob):
        u = torch.rand((n, n), generator=gen)
        a = torch.triu((u < prob).float(), diagonal=1)
        return a + a.t()
    if regular_degree is None:
        W = er(p)
    else:
        d = regular_degree
        idx = torch.arange(n)
        W = torch.zeros((n, n))
        for k in range(1, d // 2 + 1):
            W[idx, (idx + k) % n] = 1.0
            W[(idx + k) % n, idx] = 1.0
        if d % 2 == 1 and n % 2 == 0:
            W[idx, (idx + n // 2) % n] = 1.0
        perm = torch.randperm(n, generator=gen)
        W = W[perm][:, perm]
    n1 = er(noise)
    n2 = er(p * noise / (1.0 - p))
    Wn = W * (1 - n1) + (1 - W) * n2
    return adjacency_to_features(W), adjacency_to_features(Wn)


def xavier_state_dict(c_in0: int, c: int, num_blocks: int, depth: int, gen: torch.Generator,
                      c_out: Optional[int] = None, root: str = "node_embedder.ne_bm_block",
                      randomize_gn: bool = False) -> StateDict:
    """Random-init weights with the reference's shapes/keys and init law
    (xavier_uniform weights, zero bias, gn weight 1 / bias 0: models/layers.py:63-66,134-142).
    `randomize_gn` perturbs gn affine + biases so tests exercise non-trivial values."""
    c_out = c if c_out is None else c_out
    sd: StateDict = {}
    ci = c_in0
    for b in range(1, num_blocks + 1):
        co = c if b < num_blocks else c_out
        for j, cin in ((1, ci), (2, ci), (3, ci + co)):
            fin = cin
            for k in range(depth):
                bound = math.sqrt(6.0 / (fin + co))
                w = (torch.rand((co, fin, 1, 1), generator=gen) * 2 - 1) * bound
                sd[f"{root}{b}_mlp{j}.convs.{k}.weight"] = w
                bias = torch.zeros(co)
                if randomize_gn:
                    bias = (torch.rand(co, generator=gen) - 0.5) * 0.2
                sd[f"{root}{b}_mlp{j}.convs.{k}.bias"] = bias
                fin = co
            gw = torch.ones((1, co, 1, 1))
            gb = torch.zeros((1, co, 1, 1))
            if randomize_gn:
                gw = 1.0 + (torch.rand((1, co, 1, 1), generator=gen) - 0.5)
                # a few negative scales so the pooled-max sign logic is exercised
                gw = gw * torch.where(torch.rand((1, co, 1, 1), generator=gen) < 0.2, -1.0, 1.0)
                gb = (torch.rand((1, co, 1, 1), generator=gen) - 0.5) * 0.5
            sd[f"{root}{b}_mlp{j}.gn.weight"] = gw
            sd[f"{root}{b}_mlp{j}.gn.bias"] = gb
        ci = co
    return sd


# ---- emulation of the 16-bit pipeline's forward numerics (for the training-path tests) -------------------------
# The 16-bit CUDA path rounds at fixed points: stored planes, per-graph folded first-layer weights,
# hidden activations and hidden-layer weights are 16-bit; accumulation, biases, GraphNorm statistics / coefficients
# and the head are fp32.  This restates THAT forward in torch with straight-through rounding, so that autograd
# gives the exact gradient of the rounded forward.  It is the oracle of the backward kernels' arithmetic: against
# the fp32 reference gradient a 16-bit forward deviates by several 1e-2 on random-init weights whatever the backward
# does (the pooling arg-max and the loss gradient move with the forward's rounding: tools/emulate_grad_noise.py).
def _ste_round(x: Tensor, dt: torch.dtype) -> Tensor:
    return x + (x.detach().to(dt).to(x.dtype) - x.detach())


def emulated16_node_embedding(x: Tensor, sd: StateDict, dt: torch.dtype,
                              root: str = "node_embedder.ne_bm_block") -> Tensor:
    """x: (C0,n,n) one graph -> (C,n) embeddings of the 16-bit pipeline (layers.py:126-131,161-162,194-203)."""
    nb, depth = count_blocks(sd, root)
    n = x.shape[-1]
    P = n * n
    xs = _ste_round(x.reshape(x.shape[0], P), dt)
    a_prev = torch.ones(x.shape[0], dtype=x.dtype)
    s_prev = torch.zeros(x.shape[0], dtype=x.dtype)

    def mlp(inputs, pre):
        w1 = sd[f"{pre}.convs.0.weight"]
        w1 = w1.reshape(w1.shape[0], -1)
        b = sd[f"{pre}.convs.0.bias"]
        acc, off = 0, 0
        for st, a, s in inputs:
            ci = st.shape[0]
            wp = w1[:, off:off + ci]
            acc = acc + _ste_round(wp * a[None, :], dt) @ st
            b = b + wp @ s
            off += ci
        h = acc + (b[:, None] if depth > 1 else 0)
        for k in range(1, depth):
            h = _ste_round(torch.relu(h), dt)
            wk = sd[f"{pre}.convs.{k}.weight"]
            h = _ste_round(wk.reshape(wk.shape[0], -1), dt) @ h
            if k < depth - 1:
                h = h + sd[f"{pre}.convs.{k}.bias"][:, None]
        st = _ste_round(h, dt)                       # stored pre-norm planes (the last bias cancels in GraphNorm)
        mu = h.mean(1)                               # statistics come from the fp32 accumulators, not from the rounded planes
        var = (h * h).mean(1) - mu * mu
        a = sd[f"{pre}.gn.weight"].reshape(-1) / (2 * torch.sqrt(n * (var + 1e-5)))
        return st, a, sd[f"{pre}.gn.bias"].reshape(-1) - a * mu

    for i in range(1, nb + 1):
        pre = f"{root}{i}"
        y1, a1, s1 = mlp([(xs, a_prev, s_prev)], pre + "_mlp1")
        y2, a2, s2 = mlp([(xs, a_prev, s_prev)], pre + "_mlp2")
        c = y1.shape[0]
        Y1, Y2 = y1.reshape(c, n, n), y2.reshape(c, n, n)
        mult = (a1 * a2)[:, None, None] * torch.matmul(Y1, Y2) + (a1 * s2)[:, None, None] * Y1.sum(2)[:, :, None] \
            + (s1 * a2)[:, None, None] * Y2.sum(1)[:, None, :] + (s1 * s2 * n)[:, None, None]
        mult = _ste_round(mult.reshape(c, P), dt)
        ones, zeros = torch.ones(c, dtype=x.dtype), torch.zeros(c, dtype=x.dtype)
        xs, a_prev, s_prev = mlp([(mult, ones, zeros), (xs, a_prev, s_prev)], pre + "_mlp3")
    out = a_prev[:, None, None] * xs.reshape(-1, n, n) + s_prev[:, None, None]
    return out.max(-1)[0]


def emulated16_loss_and_grads(x1: Tensor, x2: Tensor, sd: StateDict, dt: torch.dtype):
    """(loss 'mean', {name: d loss / d parameter}) of the emulated 16-bit forward with an exact fp32 backward."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    total, rows = 0.0, 0
    for g in range(x1.shape[0]):
        e1 = emulated16_node_embedding(x1[g], params, dt)
        e2 = emulated16_node_embedding(x2[g], params, dt)
        s = e1.t() @ e2
        total = total + torch.nn.functional.cross_entropy(s, torch.arange(s.shape[0]), reduction="sum")
        rows += s.shape[0]
    loss = total / rows
    loss.backward()
    return float(loss.detach()), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in params.items()}
