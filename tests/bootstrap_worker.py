"""torchrun worker for tests/test_shard_protocol_cpu.py::test_host_bootstrap_callback_over_gloo: every rank calls the
bnx_allgather_fn that ShardedMap(bootstrap="host") hands to the library — through its C function pointer, with ctypes
buffers, exactly as the library does — and checks that it received every rank's 64 "handle" bytes in rank order."""
import ctypes as C
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bonxai_b200.sharded import host_allgather_callback  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cb = host_allgather_callback()
    for round_ in range(3):  # the library calls it again whenever the mailboxes are replaced
        send = (C.c_uint8 * 64)(*[(rank * 31 + round_ * 7 + k) % 251 for k in range(64)])
        recv = (C.c_uint8 * (64 * world))()
        assert cb(None, C.cast(send, C.c_void_p), C.cast(recv, C.c_void_p), 64) == 0
        for r in range(world):
            assert list(recv[64 * r:64 * r + 64]) == [(r * 31 + round_ * 7 + k) % 251 for k in range(64)], (rank, r, round_)
    dist.barrier()
    if rank == 0:
        print("BOOTSTRAP_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
