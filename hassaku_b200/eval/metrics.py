"""Batch ranking metrics behind the reference API (eval/metrics.py:4-105): dense `[B, I]` logits / y_true in,
per-user vector or batch sum out.  Top-k comes from hsk_topk_dense (ties: lower item id first — torch.topk leaves tie
order unspecified), the metric arithmetic from hsk_rank_metrics_dense; binary relevance like the reference."""
import functools

import torch

from hassaku_b200 import _C


@functools.lru_cache(maxsize=None)
def _discount_cpu(k: int) -> torch.Tensor:
    return 1. / torch.log2(torch.arange(2, k + 2).float())  # metrics.py:91 (fp32 table)


def discount_table(k: int, device) -> torch.Tensor:
    return _discount_cpu(k).to(device)


def dense_topk(logits: torch.Tensor, k: int):
    """logits.topk(k) -> (scores fp32 [B, k], ids int32 [B, k]) on the GPU kernel."""
    if not logits.is_cuda:
        raise _C.HskError('hassaku_b200 metrics need CUDA tensors (no CPU path)')
    x = logits if logits.dtype == torch.float32 else logits.float()
    if x.dim() != 2 or x.stride(1) != 1:
        x = x.reshape(x.shape[0], -1).contiguous()
    scores = torch.empty((x.shape[0], k), dtype=torch.float32, device=x.device)
    ids = torch.empty((x.shape[0], k), dtype=torch.int32, device=x.device)
    _C.topk_dense(x, k, scores, ids)
    return scores, ids


def dense_metrics(logits: torch.Tensor, y_true: torch.Tensor, ks, idx_topk: torch.Tensor = None) -> torch.Tensor:
    """-> per-user fp32 [B, len(ks), 3] (precision, recall, ndcg) for dense inputs."""
    k_max = max(ks)
    if idx_topk is None:
        _, idx_topk = dense_topk(logits, k_max)
    dev = y_true.device if y_true.is_cuda else idx_topk.device
    ids = idx_topk.to(dev, torch.int32).contiguous()
    y = y_true.to(dev, torch.float32).contiguous()
    B = ids.shape[0]
    per_user = torch.empty((B, len(ks), 3), dtype=torch.float32, device=dev)
    sums = torch.zeros((1, len(ks), 3), dtype=torch.float64, device=dev)
    counts = torch.zeros(1, dtype=torch.int64, device=dev)
    u = torch.zeros(B, dtype=torch.int64, device=dev)
    _C.rank_metrics_dense(ids, list(ks), u, y, discount_table(ids.shape[1], dev), sums, counts, per_user=per_user)
    return per_user


def _one_metric(which: int, logits, y_true, k, aggr_sum, idx_topk):
    if idx_topk is not None:
        assert idx_topk.shape[-1] == k, 'Top-k indexes have different "k" compared to the parameter function'
    res = dense_metrics(logits, y_true, [k], idx_topk)[:, 0, which]
    return res.sum() if aggr_sum else res


def recall_at_k_batch(logits: torch.Tensor, y_true: torch.Tensor, k: int = 10, aggr_sum: bool = True,
                      idx_topk: torch.Tensor = None):
    """Recall@k (metrics.py:4-36): hits / positives, 0 for users without positives."""
    return _one_metric(1, logits, y_true, k, aggr_sum, idx_topk)


def precision_at_k_batch(logits: torch.Tensor, y_true: torch.Tensor, k: int = 10, aggr_sum: bool = True,
                         idx_topk: torch.Tensor = None):
    """Precision@k (metrics.py:39-67): hits / k."""
    return _one_metric(0, logits, y_true, k, aggr_sum, idx_topk)


def ndcg_at_k_batch(logits: torch.Tensor, y_true: torch.Tensor, k: int = 10, aggr_sum: bool = True,
                    idx_topk: torch.Tensor = None):
    """NDCG@k with binary relevance (metrics.py:70-105), clamped to 1, 0 for users without positives."""
    return _one_metric(2, logits, y_true, k, aggr_sum, idx_topk)


def hit_at_k_batch(logits: torch.Tensor, y_true: torch.Tensor, k: int = 10, aggr_sum: bool = True,
                   idx_topk: torch.Tensor = None):
    """Hit@k: 1 if at least one relevant item is among the top k, else 0 (BASELINE north_star lists it; the reference has
    no such function — same signature as its three metrics).  Derived from the precision column: hits > 0."""
    res = (_one_metric(0, logits, y_true, k, False, idx_topk) > 0).float()
    return res.sum() if aggr_sum else res


def hit_from_precision(per_user: torch.Tensor) -> torch.Tensor:
    """per_user [B, n_ks, 3] (precision, recall, ndcg) -> hit [B, n_ks] in {0, 1}: precision@k > 0."""
    return (per_user[..., 0] > 0).to(per_user.dtype)


# ---- calibration distances between per-user distributions over tags / popularity buckets (eval/metrics.py:108-152) ----
# Plain torch on whatever device the inputs live on: consumers of the top-k ids, [B, n_tags] work per batch.
def hellinger_distance(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """sqrt(1/2 * sum_t (sqrt p_t - sqrt q_t)^2), shape [*]; symmetric (metrics.py:108-119)."""
    return torch.sqrt(.5 * ((torch.sqrt(p) - torch.sqrt(q)) ** 2).sum(-1))


def kl_divergence(true_p: torch.Tensor, model_q: torch.Tensor) -> torch.Tensor:
    """sum_t p_t (log p_t - log q_t), shape [*]; NaN / inf where either is 0 on an event, like the reference
    (metrics.py:122-131)."""
    return (true_p * (true_p.log() - model_q.log())).sum(-1)


def jensen_shannon_distance(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """sqrt(1/2 KL(p || m) + 1/2 KL(q || m)), m = (p + q) / 2 (metrics.py:134-152)."""
    m = .5 * (p + q)
    return torch.sqrt(.5 * (kl_divergence(p, m) + kl_divergence(q, m)))
