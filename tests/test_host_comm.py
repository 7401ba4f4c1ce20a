"""lkgpu::ShardComm (libkriging_b200/host/lkgpu_comm.hpp): the TCP exchange between the C++ host processes of a
sharded fit -- CPU only, world_size 3: the ticket queue (dynamic start queue) hands out every index exactly once
across processes and threads, ragged all-gathers arrive complete and in rank order on every rank."""
import json
import os
import socket
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SELFTEST = os.path.join(HERE, "..", "libkriging_b200", "host", "_build", "lkgpu_comm_selftest")


@pytest.mark.skipif(not os.path.isfile(SELFTEST), reason="C++ host not built (libkriging_b200/host/build_host.sh)")
def test_shard_comm_tickets_and_gather_world3():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 3
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   LKGPU_COMM_PORT_OFFSET="0")
        procs.append(subprocess.Popen([SELFTEST], stdout=subprocess.PIPE, text=True, env=env))
    outs = [json.loads(p.communicate(timeout=60)[0].strip().splitlines()[-1]) for p in procs]
    assert all(p.returncode == 0 for p in procs)
    assert [o["rank"] for o in outs] == [0, 1, 2] and all(o["ok"] and o["world"] == world for o in outs)
    assert sum(o["taken"] for o in outs) == 50
    assert sorted(o["ticket_key8"] for o in outs) == [0, 1, 2]  # a second queue starts at 0 again


@pytest.mark.skipif(not os.path.isfile(SELFTEST), reason="C++ host not built (libkriging_b200/host/build_host.sh)")
def test_shard_comm_fails_fast_when_a_process_leaves():
    """A process that dies before the exchange (a failed fit) must make the others raise, not wait for ever."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 3
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   LKGPU_COMM_PORT_OFFSET="0", LKGPU_COMM_SELFTEST_DIE="2")
        procs.append(subprocess.Popen([SELFTEST], stdout=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=60)[0].strip().splitlines()[-1] for p in procs]
    assert procs[2].returncode == 3 and "left_early" in outs[2]
    for r in (0, 1):
        assert procs[r].returncode == 1 and "connection" in json.loads(outs[r])["error"]
