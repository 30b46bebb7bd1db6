#!/usr/bin/env python
"""Follow-up to pcie_probe.py: does the KIND of page-locked memory, or who moves the bytes (copy
engine vs the synthesis kernel reading / writing mapped host memory itself), change what the
platform gives?  Run under torchrun with N ranks; every leg runs on all ranks at the same time.
Diagnostic only.  One JSON object per line on stdout (rank 0)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                      # noqa: E402
import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402
import libnyquist_b200 as nq      # noqa: E402

FRAMES = 65536 * 2


def sh(cmd):
    return subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")

    def cpu_barrier():
        if world > 1:
            dist.barrier()

    def gather(x):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    if rank == 0:
        bench.emit({"thp": sh("cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag; grep -i -E 'AnonHugePages|Hugepagesize' /proc/meminfo; cat /proc/cmdline | tr ' ' '\\n' | grep -i -E 'iommu|huge'; dmesg 2>/dev/null | grep -i -E 'iommu|DMAR' | head -5")})
    L = nq.load_library()
    nbytes = FRAMES * 7680
    n = nbytes // 4

    def host_tensor(kind):
        p = L.nq_celt_host_alloc_ex(nbytes, kind)
        assert p, kind
        arr = (C.c_float * n).from_address(p)
        t = torch.frombuffer(arr, dtype=torch.float32)
        t._keep = arr
        return t

    d_a = torch.empty(n, dtype=torch.float32, device=dev).uniform_(-100, 100)
    d_b = torch.empty(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    step = (64 << 20) // 4
    synth = nq.CeltSynth(local)
    tr_d = torch.zeros(FRAMES, dtype=torch.uint8, device=dev)
    tr_h = tr_d.cpu()

    def timeit(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        cpu_barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def report(name, dt):
        dts = gather(dt)
        cpu_barrier()
        if rank == 0:
            gb = nbytes / 1e9
            bench.emit({"leg": name, "gpus": world, "GBps_each_way_aggregate": round(world * gb / max(dts), 2),
                        "per_rank_GBps": [round(gb / d, 2) for d in dts]})

    for kind, kname in ((0, "cudaHostAlloc"), (2, "thp_registered"), (1, "write_combined")):
        try:
            ha, hb = host_tensor(kind), host_tensor(0 if kind == 1 else kind)   # (WC only for the host->device source)
        except AssertionError:
            if rank == 0:
                bench.emit({"leg": kname, "error": "allocation failed"})
            continue
        ha.copy_(d_a.cpu())
        if rank == 0:
            bench.emit({"kind": kname, "is_pinned": [bool(ha.is_pinned()), bool(hb.is_pinned())],
                        "AnonHugePages": sh("grep AnonHugePages /proc/meminfo").strip()})

        def copies(mode):
            def once():
                for o in range(0, n, step):
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s1):
                            d_a[o:o + step].copy_(ha[o:o + step], non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s2):
                            hb[o:o + step].copy_(d_b[o:o + step], non_blocking=True)
            return once
        for mode in ("h2d", "d2h", "both"):
            report(f"copy_engine_{mode}_{kname}", timeit(copies(mode)))
        # the synthesis kernel itself moving the bytes over PCIe (mapped host memory, UVA)
        pcm_d = d_b.view(-1, 2)
        coef_d = d_a.view(FRAMES, 2, 960)
        if kind != 1:
            report(f"kernel_writes_host_{kname}", timeit(lambda: synth.synth_batch_device_ptr(
                coef_d.data_ptr(), tr_d.data_ptr(), 0, 0, 0, hb.data_ptr(), 0, FRAMES, 2, 1)))
        report(f"kernel_reads_host_{kname}", timeit(lambda: synth.synth_batch_device_ptr(
            ha.data_ptr(), tr_d.data_ptr(), 0, 0, 0, pcm_d.data_ptr(), 0, FRAMES, 2, 1)))
        if kind != 1:
            report(f"kernel_reads_and_writes_host_{kname}", timeit(lambda: synth.synth_batch_device_ptr(
                ha.data_ptr(), tr_d.data_ptr(), 0, 0, 0, hb.data_ptr(), 0, FRAMES, 2, 1)))
        report(f"library_e2e_{kname}", timeit(lambda: synth.synth_batch_host_ptr(
            ha.data_ptr(), tr_h.data_ptr(), 0, hb.data_ptr(), 0, FRAMES, 2)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
