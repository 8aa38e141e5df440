# -*- coding: utf-8 -*-
"""Multi-GPU keyphrase table: documents shard over ranks, score slices join with ONE all-gather.

The reference has no distributed code; its unit of independence is the document
(one AST per text, east/relevance.py:41-47).  One process per GPU (torch.distributed, NCCL over
NVLink on the GPU box, gloo in CPU tests): every rank indexes and scores a contiguous range of
documents against the replicated keyphrases; the doc-major [D_r, K] float64 slices are padded to
a common height and joined by a single all_gather_into_tensor -- doc-major storage makes the
result a plain concatenation.  There is no other exchange on the path.
"""
import numpy as np


def partition_documents(sizes, world_size):
    """Contiguous document ranges balanced by total size (code points).

    Returns [(begin, end)] * world_size; ranges may be empty when there are fewer documents
    than ranks.  Deterministic, so every rank computes the same partition."""
    sizes = np.asarray(sizes, dtype=np.int64)
    n = len(sizes)
    bounds = [0]
    if n:
        csum = np.cumsum(sizes)
        total = int(csum[-1])
        for r in range(1, world_size):
            target = total * r / float(world_size)
            cut = int(np.searchsorted(csum, target, side="left"))
            # put the boundary document where the imbalance is smaller
            if cut < n and cut >= bounds[-1]:
                before = csum[cut - 1] if cut > 0 else 0
                if abs(csum[cut] - target) < abs(before - target):
                    cut += 1
            bounds.append(min(max(cut, bounds[-1]), n))
    else:
        bounds += [0] * (world_size - 1)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def gather_score_slices(local_scores, ranges, group=None):
    """All-gather the per-rank [D_r, K] slices into the full [D, K] table (on every rank).

    local_scores: torch float64 tensor [D_r, K] on this rank's device (CUDA for NCCL, CPU for gloo).
    ranges: the partition from partition_documents (same on all ranks)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    K = int(local_scores.shape[1])
    heights = [e - b for b, e in ranges]
    max_h = max(heights) if heights else 0
    if max_h == 0:
        return local_scores.new_zeros((0, K))
    padded = local_scores.new_zeros((max_h, K))
    padded[: local_scores.shape[0]] = local_scores
    gathered = local_scores.new_empty((world * max_h, K))
    dist.all_gather_into_tensor(gathered, padded, group=group)
    if all(h == max_h for h in heights):
        return gathered
    return torch.cat([gathered[r * max_h: r * max_h + heights[r]] for r in range(world)], dim=0)


def relevance_table_sharded(texts, prepared_keyphrases, normalized=True, device=None, group=None):
    """[D, K] float64 score table of all texts x keyphrases, computed by all ranks together.

    texts: list of raw texts (same list on every rank).  Each rank preprocesses, packs, indexes
    and scores only its own range on its own GPU; returns the gathered table as a torch tensor
    on that GPU."""
    import torch
    import torch.distributed as dist

    from east import _capi
    from east import utils
    from east.asts import utils as asts_utils

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if device is None:
        device = torch.cuda.current_device()
    ranges = partition_documents([len(t) for t in texts], world)
    begin, end = ranges[rank]
    codes, off = _capi.pack_keyphrases(prepared_keyphrases)
    K = len(prepared_keyphrases)
    if end > begin:
        cols = [utils.text_to_strings_collection(t) for t in texts[begin:end]]
        packed = [asts_utils.pack_strings_collection(c) for c in cols]
        from east.relevance import plan_batches
        table = np.empty((len(cols), K), dtype=np.float64)
        for docs in plan_batches([len(p) for p in packed]):   # small documents apart from large ones, bounded batches
            index = _capi.DeviceIndex([packed[j] for j in docs], [len(cols[j]) for j in docs], device=device)
            table[docs] = index.score_table(codes, off, normalized)
            index.close()
        local = torch.from_numpy(table).to("cuda:%d" % device)
    else:
        local = torch.zeros((0, K), dtype=torch.float64, device="cuda:%d" % device)
    return gather_score_slices(local, ranges, group)
