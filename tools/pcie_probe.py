#!/usr/bin/env python
"""Where does the host-buffer (e2e) leg's time go when several GPUs of one box are fed at once?

Run under torchrun with N ranks (one per GPU).  For k = 1, 2, 4, ... N only the first k ranks are
active while the others wait on a CPU (gloo) barrier, so ONE launch gives the whole scaling curve of
  * the platform: raw pinned cudaMemcpyAsync, host->device only, device->host only, both at once,
    with the pinned buffers allocated (a) wherever the process runs, (b) bound to the GPU's NUMA node;
  * the library: nq_celt_synth_batch_host on the same buffers, over the pipeline knobs
    NQ_HOST_SLOTS x NQ_HOST_CHUNK_MB;
  * nq_celt_synth_batch_host_multi (one process feeding k GPUs) from rank 0.
Diagnostic only; bench.py is the contract.  One JSON object per line on stdout (rank 0)."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                      # noqa: E402  (near_gpu, emit, protect_stdout)
import numpy as np                # noqa: E402
import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402
import libnyquist_b200 as nq      # noqa: E402

FRAMES = 65536 * 2                # 1 GB of coefficients in, 1 GB of samples out per rep


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=30).stdout
    except Exception as e:
        return repr(e)


def main():
    bench.protect_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")   # CPU barriers only: idle ranks must not spin on their GPU

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
        bench.emit({"topology": sh("nvidia-smi topo -m"), "lscpu": sh("lscpu | egrep -i 'model name|socket|numa|^cpu\\(s\\)|thread'"),
                    "numa_nodes_of_gpus": sh("for d in /sys/bus/pci/devices/*; do if [ \"$(cat $d/class)\" = 0x030200 ]; then echo $d $(cat $d/numa_node) $(cat $d/current_link_speed) x$(cat $d/current_link_width); fi; done"),
                    "meminfo": sh("egrep 'MemTotal|MemFree|HugePages_Total' /proc/meminfo"), "nproc": os.cpu_count(),
                    "near_gpu_cpus_rank0": sorted(bench.near_gpu_cpus(local) or [])})

    nbytes = FRAMES * 7680
    n = nbytes // 4
    bufs = {}
    for how in ("default", "near_gpu"):
        if how == "near_gpu":
            with bench.near_gpu(local):
                a = torch.empty(n, dtype=torch.float32, pin_memory=True)
                b = torch.empty(n, dtype=torch.float32, pin_memory=True)
                a.zero_(); b.zero_()
        else:
            a = torch.empty(n, dtype=torch.float32, pin_memory=True)
            b = torch.empty(n, dtype=torch.float32, pin_memory=True)
            a.zero_(); b.zero_()
        bufs[how] = (a, b)
    d_a = torch.empty(n, dtype=torch.float32, device=dev)
    d_b = torch.empty(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    step = (64 << 20) // 4

    def raw(mode, how, reps=3):
        ha, hb = bufs[how]

        def once():
            for o in range(0, n, step):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_a[o:o + step].copy_(ha[o:o + step], non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        hb[o:o + step].copy_(d_b[o:o + step], non_blocking=True)
        once()
        torch.cuda.synchronize()
        cpu_barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    synth = nq.CeltSynth(local)
    tr = torch.zeros(FRAMES, dtype=torch.uint8, pin_memory=True)
    tail = torch.empty((2, 60), dtype=torch.float32, pin_memory=True)
    bufs["default"][0].uniform_(-100, 100)
    bufs["near_gpu"][0].copy_(bufs["default"][0])

    def e2e(how, reps=3):
        ha, hb = bufs[how]

        def once():
            synth.synth_batch_host_ptr(ha.data_ptr(), tr.data_ptr(), 0, hb.data_ptr(), tail.data_ptr(), FRAMES, 2)
        once()
        cpu_barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        return (time.perf_counter() - t0) / reps

    def run(name, k, fn):
        """fn on the first k ranks at the same time; the others pass through the same barriers."""
        active = rank < k
        cpu_barrier()
        if active:
            dt = fn()
        else:
            cpu_barrier()   # the one inside fn
            dt = None
        cpu_barrier()
        dts = [d for d in gather(dt) if d is not None]
        if rank == 0:
            gb = nbytes / 1e9
            bench.emit({"leg": name, "active_gpus": k, "GBps_each_way_aggregate": round(k * gb / max(dts), 2),
                        "per_rank_GBps": [round(gb / d, 2) for d in dts]})

    ks = [k for k in (1, 2, 4, 8) if k <= world]
    for k in ks:
        for how in ("default", "near_gpu"):
            for mode in ("h2d", "d2h", "both"):
                run(f"raw_{mode}_{how}", k, lambda: raw(mode, how))
        run("e2e_default_knobs_near_gpu", k, lambda: e2e("near_gpu"))
    for slots in (1, 2, 3):
        for mb in (8, 32, 64, 256):
            os.environ["NQ_HOST_SLOTS"], os.environ["NQ_HOST_CHUNK_MB"] = str(slots), str(mb)
            run(f"e2e_slots{slots}_chunk{mb}MB_near_gpu", world, lambda: e2e("near_gpu"))
    os.environ.pop("NQ_HOST_SLOTS"); os.environ.pop("NQ_HOST_CHUNK_MB")

    # one process feeding k GPUs: rank 0 alone, everybody else idle on the CPU barrier
    import ctypes as C
    L = nq.load_library()
    for k in ks:
        cpu_barrier()
        if rank == 0:
            ha, hb = bufs["near_gpu"]
            devs = (C.c_int * k)(*range(k))

            def once():
                rc = L.nq_celt_synth_batch_host_multi(devs, k, C.c_void_p(ha.data_ptr()), C.c_void_p(tr.data_ptr()), None,
                                                      C.c_void_p(hb.data_ptr()), None, FRAMES, 2)
                assert rc == 0, rc
            once(); once()
            t0 = time.perf_counter()
            for _ in range(3):
                once()
            dt = (time.perf_counter() - t0) / 3
            bench.emit({"leg": "host_multi_one_process", "active_gpus": k, "GBps_each_way_aggregate": round(nbytes / 1e9 / dt, 2),
                        "Mframes_per_s": round(FRAMES / dt / 1e6, 3)})
        cpu_barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
