"""Multi-process host logic on CPU (gloo, world_size 2): shot sharding and the one counter all-reduce."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from surface_17_iqt_b200 import distributed

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10 ** 6, 10 ** 8 + 3):
        for w in (1, 2, 4, 8):
            parts = [distributed.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _c1) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        distributed.shard_range(10, 2, 2)


def test_no_process_group_means_no_sharding(monkeypatch, capsys):
    """Sharding and the all-reduce use one condition: under a torchrun environment WITHOUT an initialised process group
    a rank keeps the whole shot range (and says so) instead of returning a silent 1/world_size of the counters."""
    assert distributed.allreduce_counts([3, 4, 5]) == [3, 4, 5]
    assert distributed.world() == (0, 1)
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setattr(distributed, "_warned", False)
    assert distributed.world() == (0, 1)
    assert "not initialised" in capsys.readouterr().err
    assert distributed.shard_range(1000, *distributed.world()) == (0, 1000)
    assert distributed.broadcast_seed(12345678901234567890) == 12345678901234567890
    assert distributed.gather_lists([4, 5]) == [4, 5]


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {repo!r})
    import torch.distributed as dist
    from surface_17_iqt_b200 import distributed
    dist.init_process_group("gloo")
    rank, world = distributed.world()
    first, count = distributed.shard_range(1001, rank, world)
    # every rank "simulates" its shard: failed = number of shot ids divisible by 7, not0 = divisible by 3
    ids = range(first, first + count)
    local = [sum(1 for i in ids if i % 7 == 0), sum(1 for i in ids if i % 3 == 0), count]
    total = distributed.allreduce_counts(local)
    # run-til-fail sharding: one Philox key for every rank, per-rank lists concatenated in rank order
    seed = distributed.broadcast_seed(0xFEDCBA9876543210 + rank)
    runs_first, runs = distributed.shard_range(7, rank, world)
    lists = distributed.gather_lists([100 * rank + i for i in range(runs)])
    root = distributed.gather_to_root([100 * rank + i for i in range(runs)]).tolist()
    assert (rank == 0) == (len(root) == 7)
    fit = distributed.broadcast_float(0.25 if rank == 0 else None)
    assert fit == 0.25
    if rank == 0:
        assert root == lists
        print(json.dumps(total + [seed] + lists))
    dist.destroy_process_group()
""")


def test_two_rank_gloo_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("[")][-1]
    got = eval(line)
    assert got[:3] == [sum(1 for i in range(1001) if i % 7 == 0), sum(1 for i in range(1001) if i % 3 == 0), 1001]
    assert got[3] == 0xFEDCBA9876543210                      # rank 0's seed on every rank
    assert got[4:] == [0, 1, 2, 3, 100, 101, 102]           # 4 runs on rank 0, 3 on rank 1, rank order
