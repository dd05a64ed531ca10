"""One-instance-per-GPU farm: the only multi-GPU structure the path has.

Circuit instances are independent in the reference (`basic_test` proves them in a sequential loop whose body depends only
on the instance and its circuit type's setup, /root/reference/src/tests/complex_tests/mod.rs:316-410), so ranks never
exchange trace data.  `assign_instances` deals instances to ranks; `gather_proofs` brings the finished fixed-size proof
buffers to rank 0 with ONE collective (NCCL over NVLink on GPUs, gloo in the CPU tests) -- north_star: "NCCL ... only to
gather finished proofs back to rank 0".
"""
import numpy as np
import torch
import torch.distributed as dist


def assign_instances(n_instances, world_size, rank):
    """Round-robin: instance i -> rank i mod world_size (instances of one circuit type are contiguous in the scheduler's
    output, so every rank sees few distinct types and keeps few setups resident)."""
    return list(range(rank, n_instances, world_size))


def gather_proofs(local_proofs, proof_len_u64, n_instances, device=None, group=None):
    """local_proofs: {instance_index: np.uint64[proof_len_u64]} proven by this rank.  Returns on rank 0 the list of all
    n_instances proofs in instance order (None elsewhere).  One `gather` of a [slots, proof_len] int64 tensor per call."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    slots = (n_instances + world - 1) // world
    buf = torch.zeros((slots, proof_len_u64), dtype=torch.int64)
    for idx, proof in local_proofs.items():
        assert idx % world == rank, "instance proven on the wrong rank"
        buf[idx // world] = torch.from_numpy(np.ascontiguousarray(proof, dtype=np.uint64).view(np.int64))
    if world == 1:
        return [buf[i].numpy().view(np.uint64) for i in range(n_instances)]
    if device is not None:
        buf = buf.to(device, non_blocking=True)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    out = [t.cpu() for t in out]
    return [out[i % world][i // world].numpy().view(np.uint64) for i in range(n_instances)]


def bind_host_to_gpu(device_index):
    """Pins the calling thread to the CPU cores NVML reports as local to the GPU, so that the pinned host buffers this rank
    allocates afterwards (first touch) sit on the GPU's NUMA node.  With 8 ranks uploading 1.3 GB of witness per proof at the
    same time, host buffers on the wrong socket made the end-to-end step 12 % slower than on one GPU (profiles/r01_k_*).
    Returns the number of cores in the mask, or 0 when NVML is unavailable (nothing changed)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        try:   # CUDA and NVML enumerate devices differently under CUDA_VISIBLE_DEVICES: go through the UUID
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(device_index).uuid))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return 0

