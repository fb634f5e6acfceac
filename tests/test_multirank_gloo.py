"""The N > 1 host logic on CPU: two ranks over the gloo backend deal the files of a
key between them (counting.assign_files), count their share into GLOBAL sample rows
(counting.global_rows) and sum the per-rank matrices with one all-reduce -- the same
sequence counting.count_files runs on GPUs with NCCL.  The per-rank counting is done by
the C oracle here (there is no GPU in this test); the result must equal the single-rank
merge of per-file matrices (combineReadCounts) whatever the world size."""

import os
import socket

import numpy as np
import pytest

from conftest import load_golden, materialize
from tagdigger_b200 import counting, hostio

GOLD = load_golden("script.json")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, workdir, out_path):
    import torch
    import torch.distributed as dist
    from oracle import c_oracle
    os.chdir(workdir)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tags = hostio.readTags_Merged("tags.csv")
    bckeys = hostio.readBarcodeKeyfile("key.csv")
    samples, rows = counting.global_rows(bckeys)
    matrix = torch.zeros((len(samples), len(tags[1])), dtype=torch.int32)
    for f in counting.assign_files(sorted(bckeys), rank, world):
        raw = open(f, "rb").read()
        if f[-2:].lower() == "gz":
            import gzip
            raw = gzip.decompress(raw)
        m, _ = c_oracle.Counter(bckeys[f][0], tags[1], "TGCAG").count(raw)
        for k, r in enumerate(rows[f]):
            matrix[r] += torch.from_numpy(np.asarray(m[k], dtype=np.int32))
    dist.all_reduce(matrix)                      # the one collective of the path
    if rank == 0:
        np.save(out_path, matrix.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_ranks_equal_single_rank(world, tmp_path):
    import torch.multiprocessing as mp
    from oracle import c_oracle
    materialize(GOLD["filesets"]["lanes"], tmp_path)
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path), out), nprocs=world, join=True)
    got = np.load(out)

    old = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        tags = hostio.readTags_Merged("tags.csv")
        bckeys = hostio.readBarcodeKeyfile("key.csv")
        per_file = {}
        for f in bckeys:
            raw = open(f, "rb").read()
            if f[-2:].lower() == "gz":
                import gzip
                raw = gzip.decompress(raw)
            per_file[f] = c_oracle.Counter(bckeys[f][0], tags[1], "TGCAG").count(raw)[0].tolist()
        want = hostio.combineReadCounts(per_file, bckeys)
    finally:
        os.chdir(old)
    assert got.tolist() == want[1]
