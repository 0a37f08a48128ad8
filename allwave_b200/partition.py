"""Host-side greedy partitioner of the pair list across the GPUs of one box.

The path shards by independent pairs: every GPU holds the whole sequence store and
aligns its own shard; only results are gathered, so there is no collective.  Pairs are sorted by
predicted cost and assigned longest-processing-time-first to the least-loaded GPU.
Predicted cost of a pair = (max(len_q, len_t) * d)^2 ~ wavefront cells (s^2 with s ~ len * d),
with d the estimated divergence (mash distance when supplied, otherwise a constant).
"""
import heapq


def predicted_cost(len_q, len_t, divergence=None):
    d = 0.05 if divergence is None else min(0.5, max(divergence, 1e-3))
    s = max(len_q, len_t) * d + abs(len_q - len_t)
    return s * s + (len_q + len_t)


def partition_pairs(pairs, lens, n_parts, divergence=None):
    """pairs: list of (q, t); lens: sequence lengths; divergence: optional dict {(q,t): d} or list aligned with pairs
    (Context.estimate_divergence).  Returns n_parts lists; every pair appears exactly once; deterministic."""
    if n_parts <= 1:
        return [list(pairs)]
    costed = []
    for idx, (q, t) in enumerate(pairs):
        if divergence is None:
            d = None
        elif isinstance(divergence, dict):
            d = divergence.get((q, t))
        else:
            d = divergence[idx]
        costed.append((predicted_cost(lens[q], lens[t], d), idx))
    costed.sort(key=lambda x: (-x[0], x[1]))
    heap = [(0.0, r) for r in range(n_parts)]
    heapq.heapify(heap)
    shards = [[] for _ in range(n_parts)]
    for cost, idx in costed:
        load, r = heapq.heappop(heap)
        shards[r].append(idx)
        heapq.heappush(heap, (load + cost, r))
    return [[pairs[i] for i in sorted(s)] for s in shards]


def shard_loads(shards, lens, divergence=None):
    return [sum(predicted_cost(lens[q], lens[t], divergence.get((q, t)) if divergence else None) for q, t in s) for s in shards]
