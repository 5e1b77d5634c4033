#!/usr/bin/env python
"""Random-read rate from another rank's device buffer mapped through CUDA IPC (kmer_b200_peer_buffer_*), next to the
same reads from local memory: what the peer-positions search pays per foreign candidate list. Run under torchrun."""
import os

import torch
import torch.distributed as dist

import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kmer_index_b200 as kb  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_bytes = int(os.environ.get("PROBE_BYTES", 4 << 30))
    mine, handle = kb.peer_buffer_create(local, n_bytes)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    other = kb.peer_buffer_open(local, handles[(rank + 1) % world])
    torch.cuda.synchronize()
    dist.barrier()
    here = kb.gather_probe_at(mine, n_bytes)
    there = kb.gather_probe_at(other, n_bytes)
    print(f"rank {rank}: local {here / 1e9:.2f} G reads/s, peer (IPC mapping of rank {(rank + 1) % world}) {there / 1e9:.2f} G reads/s",
          flush=True)
    dist.barrier()
    kb.peer_buffer_release(local, other, True)
    dist.barrier()
    kb.peer_buffer_release(local, mine, False)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
