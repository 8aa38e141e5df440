"""CPU, world_size 2, gloo: host-side sharding logic of the multi-GPU table.

The per-rank score slices are produced by the CPU oracle here (the product path needs a GPU);
what is under test is the partition and the single all-gather that joins doc-major slices."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_docs, out_dir):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import synth
        from east import _capi, distributed, utils
        from oracle import oracle
        docs = [synth.document(1500 + 700 * (j % 3), 50 + j) for j in range(n_docs)]
        kps = [utils.prepare_text(k) for k in synth.keyphrases(7)]
        codes, off = _capi.pack_keyphrases(kps)
        ranges = distributed.partition_documents([len(d) for d in docs], world)
        b, e = ranges[rank]
        rows = [oracle.OracleEASA(utils.text_to_strings_collection(d)).score_many(codes, off, True) for d in docs[b:e]]
        local = torch.from_numpy(np.array(rows, dtype=np.float64).reshape(e - b, len(kps)))
        full = distributed.gather_score_slices(local, ranges)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_docs", [5, 1])
def test_sharded_table_equals_single_process(tmp_path, oracle_mod, n_docs):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_docs, str(tmp_path)), nprocs=world, join=True)
    import synth
    from east import _capi, utils
    docs = [synth.document(1500 + 700 * (j % 3), 50 + j) for j in range(n_docs)]
    kps = [utils.prepare_text(k) for k in synth.keyphrases(7)]
    codes, off = _capi.pack_keyphrases(kps)
    expect = np.array([oracle_mod.OracleEASA(utils.text_to_strings_collection(d)).score_many(codes, off, True)
                       for d in docs])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert got.shape == expect.shape
        assert np.array_equal(got.view(np.uint64), expect.view(np.uint64))  # pure concatenation: bit-identical


def test_partition_documents_properties():
    from east import distributed
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(0, 40))
        world = int(rng.integers(1, 9))
        sizes = rng.integers(1, 1000, size=n)
        parts = distributed.partition_documents(sizes, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert all(b <= e for b, e in parts)
        if n >= 4 * world:
            loads = [int(sizes[b:e].sum()) for b, e in parts]
            assert max(loads) <= sizes.sum() / world + sizes.max()
    assert distributed.partition_documents([10, 10, 10, 10], 2) == [(0, 2), (2, 4)]
    assert distributed.partition_documents([], 3) == [(0, 0), (0, 0), (0, 0)]
