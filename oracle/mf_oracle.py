"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — torch-CPU restatement of the reference hot path.

Every function names the reference lines it follows (paths relative to karapostK/hassaku).  The
arithmetic lives in torch (the reference's own dependency): the restatement issues the same ATen
op sequence the reference's Python issues, so on the same torch build it reproduces the reference
to the last bit on CPU (checked in tests/test_oracle_golden.py against fixtures generated from the
real reference by oracle/make_golden.py).

Pinned: yes — metrics by framework_tests/eval/test_metrics.py:29-69; the rest by reference-generated
fixtures under tests/golden/.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch
from torch import nn

K_VALUES = [5, 10, 50, 100]  # eval/eval.py:20


# --------------------------------------------------------------------------------------------
# a1-a4  forward  (algorithms/base_classes.py:99-108, algorithms/sgd_alg.py:148-179)
# --------------------------------------------------------------------------------------------
class OracleMF(nn.Module):
    """Restates SGDMatrixFactorization (algorithms/sgd_alg.py:110-184).  Same parameter names so a
    reference `state_dict` loads directly (base_classes.py:156-165)."""

    def __init__(self, n_users: int, n_items: int, embedding_dim: int = 100, use_user_bias: bool = False,
                 use_item_bias: bool = False, use_global_bias: bool = False):
        super().__init__()
        self.n_users, self.n_items, self.embedding_dim = n_users, n_items, embedding_dim
        self.use_user_bias, self.use_item_bias, self.use_global_bias = use_user_bias, use_item_bias, use_global_bias
        self.user_embeddings = nn.Embedding(n_users, embedding_dim)  # sgd_alg.py:127
        self.item_embeddings = nn.Embedding(n_items, embedding_dim)  # sgd_alg.py:128
        if use_user_bias:
            self.user_bias = nn.Embedding(n_users, 1)  # sgd_alg.py:131
        if use_item_bias:
            self.item_bias = nn.Embedding(n_items, 1)  # sgd_alg.py:133
        # train/utils.py:11-13: every nn.Embedding ~ N(0, (0.1 / weight.shape[-1])^2)
        for m in self.modules():
            if isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, std=.1 / m.weight.shape[-1])
        if use_global_bias:
            self.global_bias = nn.Parameter(torch.zeros(1))  # sgd_alg.py:137-138

    def get_user_representations(self, u_idxs):  # sgd_alg.py:148-152
        if self.use_user_bias:
            return self.user_embeddings(u_idxs), self.user_bias(u_idxs)
        return self.user_embeddings(u_idxs)

    def get_item_representations(self, i_idxs):  # sgd_alg.py:154-157
        if self.use_item_bias:
            return self.item_embeddings(i_idxs), self.item_bias(i_idxs).squeeze()
        return self.item_embeddings(i_idxs)

    def combine_user_item_representations(self, u_repr, i_repr):  # sgd_alg.py:159-179
        u_embed, u_bias = u_repr if isinstance(u_repr, tuple) else (u_repr, None)
        i_embed, i_bias = i_repr if isinstance(i_repr, tuple) else (i_repr, None)
        out = (u_embed[:, None, :] * i_embed).sum(dim=-1)  # sgd_alg.py:171 (materialises [B,N+1,d])
        if self.use_user_bias:
            out += u_bias
        if self.use_item_bias:
            out += i_bias
        if self.use_global_bias:
            out += self.global_bias
        return out

    def forward(self, u_idxs, i_idxs):  # base_classes.py:99-108
        return self.combine_user_item_representations(self.get_user_representations(u_idxs),
                                                      self.get_item_representations(i_idxs))


# --------------------------------------------------------------------------------------------
# a5-a6  losses  (train/rec_losses.py)
# --------------------------------------------------------------------------------------------
def bce_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """rec_losses.py:39-53."""
    return nn.BCEWithLogitsLoss()(logits.flatten(), labels.flatten())


def bpr_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """rec_losses.py:68-88.  `labels` is float64 in the reference loader (data/dataloader.py:127) which
    promotes the result to float64."""
    pos_logits = logits[:, 0].unsqueeze(1)
    neg_logits = logits[:, 1:]
    lab = torch.repeat_interleave(labels[:, 0], neg_logits.shape[1])
    diff = pos_logits - neg_logits
    return nn.BCEWithLogitsLoss()(diff.flatten(), lab.flatten())


def sampled_softmax_loss(logits: torch.Tensor, n_items: int, neg_train: int,
                         train_neg_strategy: str = 'uniform') -> torch.Tensor:
    """rec_losses.py:117-139 (mutates `logits[:, 1:]` in place exactly like the reference, :134)."""
    pos_logits_sum = - logits[:, 0]
    if train_neg_strategy == 'uniform':
        logits[:, 1:] += math.log(n_items / neg_train)
    return (pos_logits_sum + torch.logsumexp(logits, dim=-1)).mean()


def compute_loss(kind: str, logits, labels, n_items=None, neg_train=None, strategy='uniform'):
    if kind == 'bpr':
        return bpr_loss(logits, labels)
    if kind == 'sampled_softmax':
        return sampled_softmax_loss(logits, n_items, neg_train, strategy)
    if kind == 'bce':
        return bce_loss(logits, labels)
    raise ValueError(kind)


def make_labels(B: int, N1: int) -> torch.Tensor:
    """data/dataloader.py:126-128: float64 labels, column 0 = 1."""
    lab = torch.zeros(B, N1, dtype=torch.float64)
    lab[:, 0] = 1.
    return lab


# --------------------------------------------------------------------------------------------
# a7-a9  one training step  (train/trainer.py:128-148) with torch.optim.AdamW (trainer.py:52-53)
# --------------------------------------------------------------------------------------------
class OracleTrainer:
    """The reference's per-batch body (train/trainer.py:133-148) around an OracleMF (or the real
    reference model — anything with the same forward)."""

    def __init__(self, model: nn.Module, loss_kind: str, lr: float, wd: float, optimizer: str = 'adamw',
                 neg_train: Optional[int] = None, strategy: str = 'uniform'):
        self.model = model
        self.loss_kind, self.neg_train, self.strategy = loss_kind, neg_train, strategy
        if optimizer == 'adam':  # trainer.py:48-55
            self.optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
        elif optimizer == 'adagrad':
            self.optimizer = torch.optim.Adagrad(model.parameters(), lr=lr, weight_decay=wd)
        elif optimizer == 'adamw':
            self.optimizer = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd)
        else:
            raise ValueError(f"Optimizer {optimizer} not yet implemented")

    def step(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor, labels: Optional[torch.Tensor] = None,
             keep: bool = False) -> Dict[str, torch.Tensor]:
        if labels is None:
            labels = make_labels(*i_idxs.shape).to(i_idxs.device)
        out = self.model(u_idxs, i_idxs)  # trainer.py:133
        scores = out.detach().clone() if keep else None
        cap = {}
        if keep:  # a hook registered before the in-place shift of sampled-softmax sees dL/d(model output)
            out.register_hook(lambda gr: cap.__setitem__('g', gr.detach().clone()))
        loss = compute_loss(self.loss_kind, out, labels, getattr(self.model, 'n_items', None), self.neg_train,
                            self.strategy)  # trainer.py:135
        loss.backward()  # trainer.py:146
        res = {'loss': loss.detach().clone()}
        if keep:
            res['scores'] = scores
            res['dscores'] = cap['g']
            res['grads'] = {n: p.grad.detach().clone() for n, p in self.model.named_parameters()}
        self.optimizer.step()  # trainer.py:147
        self.optimizer.zero_grad()  # trainer.py:148
        return res


def closed_form_dscores(scores: np.ndarray, loss_kind: str, n_items: int = 0, neg_train: int = 0) -> np.ndarray:
    """SURVEY Appendix A.2-A.3 — the analytic dL/dscores in float64 (an independent cross-check of the
    autograd path above; never the gate by itself)."""
    s = scores.astype(np.float64)
    B, N1 = s.shape
    N = N1 - 1
    ds = np.zeros_like(s)
    if loss_kind == 'bpr':
        x = s[:, :1] - s[:, 1:]
        dx = (1. / (1. + np.exp(-x)) - 1.) / (B * N)
        ds[:, 1:] = -dx
        ds[:, 0] = dx.sum(1)
    elif loss_kind == 'sampled_softmax':
        s2 = s.copy()
        s2[:, 1:] += math.log(n_items / neg_train)
        e = np.exp(s2 - s2.max(1, keepdims=True))
        ds = e / e.sum(1, keepdims=True)
        ds[:, 0] -= 1.
        ds /= B
    elif loss_kind == 'bce':
        y = np.zeros_like(s)
        y[:, 0] = 1.
        ds = (1. / (1. + np.exp(-s)) - y) / (B * N1)
    else:
        raise ValueError(loss_kind)
    return ds


# --------------------------------------------------------------------------------------------
# a10  negative sampling with the reference's semantics (data/dataloader.py:56-57, 92-129)
# --------------------------------------------------------------------------------------------
def sample_negatives_reference(u_idxs: np.ndarray, pos_idxs: np.ndarray, n_neg: int, n_items: int, indptr: np.ndarray,
                               indices: np.ndarray, rng: np.random.RandomState):
    """Same loop as `_neg_sampling_collate_fn`: redraw every entry that `np.isin(row, train_row,
    assume_unique=True)` flags, until none is flagged."""
    B = len(u_idxs)
    neg = np.empty((B, n_neg), dtype=np.int64)
    mask = np.ones_like(neg, dtype=bool)
    to_resample = mask.sum()
    while True:
        neg[mask] = rng.randint(0, high=n_items, size=to_resample)
        for i in range(B):
            row = indices[indptr[u_idxs[i]]:indptr[u_idxs[i] + 1]]
            mask[i] = np.isin(neg[i], row, assume_unique=True)
        to_resample = mask.sum()
        if to_resample == 0:
            break
    i_idxs = np.column_stack([pos_idxs, neg]).astype(np.int64)
    labels = np.zeros_like(i_idxs, dtype=float)
    labels[:, :1] = 1.
    return u_idxs.astype(np.int64), i_idxs, labels


# --------------------------------------------------------------------------------------------
# a12-a17  full-rank evaluation  (eval/eval.py:54-99, 101-118, 237-253; eval/metrics.py:4-105)
# --------------------------------------------------------------------------------------------
def recall_at_k(y_true: torch.Tensor, idx_topk: torch.Tensor) -> torch.Tensor:
    """metrics.py:4-36 (per-user vector, aggr_sum=False)."""
    col = torch.arange(y_true.shape[0]).unsqueeze(-1)
    num = y_true[col, idx_topk].sum(dim=-1)
    den = y_true.sum(dim=-1)
    recall = num / den
    recall[torch.isnan(recall)] = .0
    return recall


def precision_at_k(y_true: torch.Tensor, idx_topk: torch.Tensor) -> torch.Tensor:
    """metrics.py:39-67."""
    col = torch.arange(y_true.shape[0]).unsqueeze(-1)
    return y_true[col, idx_topk].sum(dim=-1) / idx_topk.shape[-1]


def ndcg_at_k(y_true: torch.Tensor, idx_topk: torch.Tensor) -> torch.Tensor:
    """metrics.py:70-105."""
    k = idx_topk.shape[-1]
    col = torch.arange(y_true.shape[0]).unsqueeze(-1)
    discount = 1. / torch.log2(torch.arange(2, k + 2).float())
    dcg = (y_true[col, idx_topk] * discount).sum(-1)
    idcg = (y_true.topk(k).values * discount).sum(-1)
    ndcg = dcg / idcg
    ndcg[torch.isnan(ndcg)] = .0
    return ndcg.clamp(max=1.)


METRIC_FUNCS = (('precision@{}', precision_at_k), ('recall@{}', recall_at_k), ('ndcg@{}', ndcg_at_k))


def masked_scores(model: OracleMF, u_idxs: torch.Tensor, exclude_csr) -> torch.Tensor:
    """eval/eval.py:237-251: item representations once, combine, -inf on exclude_data[u]."""
    with torch.no_grad():
        i_repr = model.get_item_representations(torch.arange(model.n_items))
        out = model.combine_user_item_representations(model.get_user_representations(u_idxs), i_repr)
        mask = torch.tensor(exclude_csr[u_idxs.numpy()].toarray().astype(bool))
        out[mask] = -torch.inf
    return out


class OracleFullEvaluator:
    """eval/eval.py:14-118 with aggr_by_group=True: per batch fp32 `.sum().item()` into python floats,
    divided by the number of users seen (users without positives included)."""

    def __init__(self, n_groups: int = 0, user_to_user_group: Optional[torch.Tensor] = None,
                 k_values: Sequence[int] = tuple(K_VALUES)):
        self.n_groups, self.u2g, self.k_values = n_groups, user_to_user_group, list(k_values)
        self.sums: Dict[int, Dict[str, float]] = {}
        self.n_entries: Dict[int, int] = {}

    def eval_batch(self, u_idxs: torch.Tensor, logits: torch.Tensor, y_true: torch.Tensor,
                   idx_topk: Optional[torch.Tensor] = None):
        ks = sorted(self.k_values, reverse=True)
        if idx_topk is None:
            idx_topk = logits.topk(k=ks[0]).indices  # eval.py:63
        self.n_entries[-1] = self.n_entries.get(-1, 0) + len(u_idxs)
        grp = self.u2g[u_idxs] if self.n_groups > 0 else None
        for g in range(self.n_groups):
            self.n_entries[g] = self.n_entries.get(g, 0) + int((grp == g).sum())
        for k in ks:
            idx_topk = idx_topk[:, :k]
            for name, fn in METRIC_FUNCS:
                res = fn(y_true, idx_topk)
                d = self.sums.setdefault(-1, {})
                d[name.format(k)] = d.get(name.format(k), 0) + res.sum().item()
                for g in range(self.n_groups):
                    d = self.sums.setdefault(g, {})
                    d[name.format(k)] = d.get(name.format(k), 0) + res[grp == g].sum().item()

    def get_results(self) -> Dict[str, float]:  # eval.py:101-118
        out = {}
        for g, d in self.sums.items():
            for name, v in d.items():
                out[name if g == -1 else f'group_{g}_{name}'] = v / self.n_entries[g]
        self.sums, self.n_entries = {}, {}
        return out


def evaluate(model: OracleMF, labels_csr, exclude_csr, eval_batch_size: int = 256, n_groups: int = 0,
             user_to_user_group=None, users: Optional[np.ndarray] = None, return_topk: bool = False):
    """evaluate_recommender_algorithm, SGD branch (eval/eval.py:237-255) fed the way FullEvalDataset
    feeds it (data/dataset.py:199-201: dense float32 label rows)."""
    ev = OracleFullEvaluator(n_groups, user_to_user_group)
    users = np.arange(model.n_users) if users is None else users
    topk_all = []
    for s in range(0, len(users), eval_batch_size):
        u = torch.from_numpy(users[s:s + eval_batch_size].astype(np.int64))
        out = masked_scores(model, u, exclude_csr)
        y_true = torch.from_numpy(labels_csr[u.numpy()].toarray().astype('float32'))
        idx = out.topk(k=max(ev.k_values)).indices
        ev.eval_batch(u, out, y_true, idx)
        if return_topk:
            topk_all.append(idx)
    res = ev.get_results()
    if return_topk:
        return res, torch.cat(topk_all)
    return res


def metrics_from_topk(topk_ids: np.ndarray, users: np.ndarray, labels_csr, k_values=tuple(K_VALUES)):
    """SURVEY A.7 — per-user metrics from top-k ids + CSR labels (no dense rows).  Returns
    {name: float64 per-user vector}; used to check `hsk_rank_metrics` at sizes where dense `[B,I]`
    label rows cannot be built."""
    kmax = topk_ids.shape[1]
    w = (1. / torch.log2(torch.arange(2, kmax + 2).float())).numpy().astype(np.float32)
    cw = np.concatenate([[0.], np.cumsum(w.astype(np.float64))])
    out = {f'{m}@{k}': np.zeros(len(users)) for k in k_values for m in ('precision', 'recall', 'ndcg')}
    indptr, indices = labels_csr.indptr, labels_csr.indices
    for r, u in enumerate(users):
        pos = indices[indptr[u]:indptr[u + 1]]
        npos = len(pos)
        hit = np.isin(topk_ids[r], pos).astype(np.float64)
        for k in k_values:
            h = hit[:k]
            out[f'precision@{k}'][r] = h.sum() / k
            out[f'recall@{k}'][r] = h.sum() / npos if npos > 0 else 0.
            out[f'ndcg@{k}'][r] = min(1., (h * w[:k]).sum() / cw[min(k, npos)]) if npos > 0 else 0.
    return out


# --------------------------------------------------------------------------------------------
# row-sparse "lazy" AdamW — NOT in the reference (which always runs dense torch.optim.AdamW); this restates the
# documented semantics of hsk_adamw_rows_lazy (= torch.optim.SparseAdam's masked update with the global step count, plus
# decoupled weight decay on the touched rows) so that the CUDA kernel has a checker.
# --------------------------------------------------------------------------------------------
class OracleLazyAdamW:
    def __init__(self, model: nn.Module, lr: float, wd: float, betas=(0.9, 0.999), eps: float = 1e-8):
        self.model, self.lr, self.wd, self.betas, self.eps, self.t = model, lr, wd, betas, eps, 0
        self.state = {n: (torch.zeros_like(p), torch.zeros_like(p)) for n, p in model.named_parameters()}

    @torch.no_grad()
    def step(self, u_idxs: torch.Tensor, i_idxs: torch.Tensor):
        """Consumes p.grad of every parameter; rows not indexed by this batch are left untouched."""
        self.t += 1
        b1, b2 = self.betas
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        rows_u, rows_i = torch.unique(u_idxs), torch.unique(i_idxs)
        for n, p in self.model.named_parameters():
            m, v = self.state[n]
            if n.startswith('user_'):
                sel = rows_u
            elif n.startswith('item_'):
                sel = rows_i
            else:
                sel = torch.arange(p.shape[0])
            g = p.grad[sel]
            pp = p[sel] * (1 - self.lr * self.wd)
            mm = torch.lerp(m[sel], g, 1 - b1)
            vv = v[sel] * b2 + (1 - b2) * g * g
            denom = vv.sqrt() / math.sqrt(bc2) + self.eps
            pp = pp - (self.lr / bc1) * mm / denom
            p[sel], m[sel], v[sel] = pp, mm, vv
            p.grad = None
