#!/usr/bin/env python
"""bench.py -- EP-head train tokens/s (fwd + bwd + all-reduce + LARS) and % of the HBM roofline.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (the reference algorithm on the host cores)

Workload (BASELINE.json configs[1]): EP head on DINOv2 ViT-L/14 tokens, N=257, D=1024, bf16, per-GPU
batch 1024, 1000 classes, M = 32 queries (the reference default, main_linprobe.py:113); synthetic N(0,1)
tokens rounded to bf16, random-init head.  A "step" is one full optimisation step of the head on one
batch that is already resident in HBM; batches cycle through a pool of `--pool` distinct batches
(8 x 539 MB >> 126 MB L2, so every step streams its tokens from HBM).  Weak scaling: per-GPU batch fixed.

One JSON line on stdout (rank 0).  Extra keys next to the contract's:
  roofline      the slower of the token-streaming kernels (the one-pass "fused fwd" / "fused bwd" where the
                shape allows them, else ks<2> logits+softmax, kp<0> pool, ks<1> dS, kp<1> dq), each timed
                with CUDA events on its launching stream; achieved = B*N*D*2 bytes (algorithmic: each bf16
                token crosses HBM once per direction) / duration; peak = MEASURED_PEAKS.json
                hbm_gbs.  roofline_all_streaming_kernels lists all of them; `traffic` is the kernel's
                measured DRAM bytes (ncu, profiles/traffic.json).
  step_roofline_frac   (2*B*N*D*2 bytes / step time) / peak -- the whole step against the roofline.
  cpu_baseline  the oracle (CPU restatement of the reference head, torch fp32, all host threads) on a
                bounded sample of the same workload.
  e2e           the same step through EPHeadTrainer.train_step_host: pinned-host tokens and labels copied
                H2D and the loss read back D2H inside the timed region, every step; h2d_gbs_per_gpu is
                the copy rate each GPU saw.  Each rank pins itself to its GPU's NUMA-local cores first
                (NVML's ideal CPU affinity) so that the pinned staging buffers are allocated there.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {   # BASELINE.json configs; per-GPU batch 1024 except c1
    "c1": dict(B=64, N=197, D=768, K=1000, name="MAE ViT-B/16 tokens N=197 D=768"),
    "c2": dict(B=1024, N=257, D=1024, K=1000, name="DINOv2 ViT-L/14 tokens N=257 D=1024"),
    "c3": dict(B=1024, N=256, D=1152, K=1000, name="SigLIP2 SO400M/14 tokens N=256 D=1152"),
    "c4": dict(B=1024, N=730, D=1664, K=1000, name="MetaCLIP2 bigG/14-378 tokens N=730 D=1664"),
    "c5": dict(B=1024, N=201, D=4096, K=1000, name="DINOv3 ViT-7B/16 tokens N=201 D=4096"),
}
METRIC = "EP-head train tokens/sec (fwd+bwd)"
UNIT = "tokens/s"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pin_to_gpu_numa(index):
    """Bind this process to the CPUs NVML reports as local to GPU `index`; returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].strip().isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"{len(cpus)} cpus {cpus[0]}-{cpus[-1]}"
    except Exception as e:                                        # a report, never a gate
        return f"unchanged ({type(e).__name__})"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_step(cfg, M, sample_B, iters, warmup, threads=None):
    """The reference head (oracle restatement: value projection on every token, autograd backward, LARS)
    on the host cores; returns (tokens/s, seconds per step, cores)."""
    import torch
    from oracle import ep_oracle as O
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    N, D, K = cfg["N"], cfg["D"], cfg["K"]
    p = O.build_head(D, M, K, seed=0)
    x = O.synthetic_tokens(sample_B, N, D, seed=1234).float()
    y = O.synthetic_labels(sample_B, K)
    mus = None
    ts = []
    for it in range(warmup + iters):
        t0 = time.perf_counter()
        r = O.head_loss_and_grads(p, x, y, dtype=torch.float32)
        names = [n for n, _ in p.trainable()]
        params = [t.detach() for _, t in p.trainable()]
        mus = mus or [torch.zeros_like(t) for t in params]
        new_p, mus = O.lars_step(params, [r["grad." + n] for n in names], mus, lr=0.1)
        p.cls_token, p.v_weight, p.fc_weight, p.fc_bias = new_p[0], new_p[1], new_p[-2], new_p[-1]
        dt = time.perf_counter() - t0
        if it >= warmup:
            ts.append(dt)
    sec = sum(ts) / len(ts)
    return sample_B * N / sec, sec, cores


def config_dict(args, cfg, world, comm_sms=None, extra=None):
    """`config` of the JSON line -- the same keys and values for our arm and the reference arm."""
    B, N, D, K, M = cfg["B"], cfg["N"], cfg["D"], cfg["K"], args.queries
    c = {"workload": f"EP head (M={M}) on {cfg['name']}, per-GPU batch {B}, {K} classes, fwd+bwd+allreduce+LARS",
         "per_gpu_batch": B, "global_batch": B * world, "tokens": N, "dim": D, "queries": M, "classes": K,
         "parallelism": f"dp{world}"}
    c.update(extra or {})
    return c


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the same step as our arm: the config's full per-GPU batch on the host cores (about 0.8 s per step for c2 on a
    # 2-socket host); --cpu-sample bounds it for the largest configs
    sample_B = args.cpu_sample or cfg["B"]
    val, sec, cores = cpu_reference_step(cfg, args.queries, sample_B, args.steps, args.warmup)
    sample = (f"{sample_B} of {cfg['B']} samples per step of the same workload, reference formulation "
              f"(value projection on every token) fwd+bwd+LARS, torch fp32, {cores} threads; rank 0 only")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, cfg, int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def time_kernel(fn, iters, flush):
    """Average device time of fn() in ms: CUDA events on the launching (current) stream, one launch per
    measurement, the L2 flushed by rotating inputs (fn takes the iteration index)."""
    import torch
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        flush(i)
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--queries", type=int, default=32)
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--pool", type=int, default=8, help="distinct resident batches cycled through")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="samples per CPU step (0 = reference arm: the full per-GPU batch; cpu_baseline leg: 64)")
    ap.add_argument("--comm-sms", type=int, default=-1,
                    help="SMs reserved for the overlapped all-reduce (N > 1); -1 = the trainer's default for N")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-mode", type=int, default=0, help="0 auto, 1 general kernels, 2 tcgen05 only")
    ap.add_argument("--debug-bits", type=int, default=0, help="ep_set_debug() developer knobs for the whole run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["B"] = args.batch
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch
    import torch.distributed as dist
    import efficient_probing_b200 as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: efficient_probing_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = pin_to_gpu_numa(local)                            # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the gradient all-reduce overlaps the token-streaming backward kernels, which leave `comm_sms` SMs free
        # for it (EPHeadTrainer): keep the collective within them
        if args.comm_sms < 0:
            args.comm_sms = E.trainer.default_comm_sms(world)
        if args.comm_sms > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.comm_sms))
        dist.init_process_group("nccl", device_id=dev)
    B, N, D, K, M = cfg["B"], cfg["N"], cfg["D"], cfg["K"], args.queries
    lib = E._lib.load()
    lib.ep_set_kernel_mode(args.kernel_mode)
    lib.ep_set_debug(args.debug_bits)

    torch.manual_seed(0)                                         # identical init on every rank (DDP broadcast)
    head = E.make_ep_head(D, M, K).to(dev)
    tr = E.EPHeadTrainer(head, B, N, lr=0.1, use_graph=not args.no_graph, comm_sms=args.comm_sms)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)   # data seed = seed + rank (main_linprobe.py:517)
    pool_x, pool_y = [], []
    for i in range(args.pool):
        x = torch.randn(B, N, D, device=dev, generator=gen, dtype=torch.float32).to(torch.bfloat16)
        y = torch.randint(0, K, (B,), device=dev, generator=gen)
        pool_x.append(x); pool_y.append(y)
        tr.register_batch(x, y)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput: W warm-up + exactly K timed steps ----------------
    for i in range(args.pool):                                   # set-up: capture every pool slot's graph (no training)
        tr.prepare(pool_x[i], pool_y[i])
    for i in range(args.warmup):
        tr.train_step(pool_x[i % args.pool], pool_y[i % args.pool])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        tr.train_step(pool_x[i % args.pool], pool_y[i % args.pool])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = B * N * world * args.steps / (total_ms * 1e-3)
    loss = tr.mean_loss()
    # replicas must still be identical: max |p - p_rank0| over all parameters and ranks
    replica_diff = None
    if world > 1:
        dmax = torch.zeros(1, device=dev)
        for prm in tr.params:
            ref = prm.detach().clone()
            dist.broadcast(ref, src=0)
            dmax = torch.maximum(dmax, (prm.detach() - ref).abs().max().reshape(1))
        dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
        replica_diff = float(dmax)

    # ---------------- end to end from pinned host buffers ----------------
    hx = [pool_x[i].cpu().pin_memory() for i in range(2)]
    hy = [pool_y[i].cpu().pin_memory() for i in range(2)]
    for i in range(3):
        tr.train_step_host(hx[i % 2], hy[i % 2], next_x_host=hx[(i + 1) % 2], next_targets_host=hy[(i + 1) % 2])
    barrier()
    e2e_steps = args.steps
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(e2e_steps):               # step i+1's H2D copy is issued while step i computes
        tr.train_step_host(hx[i % 2], hy[i % 2], next_x_host=hx[(i + 1) % 2], next_targets_host=hy[(i + 1) % 2])
    f1.record()
    barrier()
    ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = B * N * world * e2e_steps / (float(ms2) * 1e-3)
    h2d = B * N * D * 2 + B * 8
    d2h = 4

    # ---------------- the token-streaming kernels alone (rank 0's GPU), for the roofline ----------------
    # ep_fwd / ep_bwd_pool are replayed outside the graph with the library's per-kernel CUDA-event timers on
    # (ep_set_debug(32): events recorded on the launching stream around every kernel of the call).  Each of
    # the four streaming kernels reads the bf16 tokens of the batch exactly once, so its algorithmic bytes
    # per launch are B*N*D*2 (2*D bytes per token per pass).
    peak, peak_src = peaks()
    alg_bytes = B * N * D * 2
    pool_mod, s = head[0], E._lib.stream_ptr(dev)
    xt = E._lib.x_dtype_code(pool_x[0])
    lib.ep_set_debug(32 | args.debug_bits)
    E._lib.kernel_timings()
    reps = 6
    for i in range(reps + 2):
        x = pool_x[i % args.pool]                                  # rotating the pool is the L2 flush
        E._lib.check(lib.ep_fwd(x.data_ptr(), xt, pool_mod.cls_token.data_ptr(), pool_mod.v.weight.data_ptr(), None,
                                float(pool_mod.scale), B, N, D, M, 1, tr.out.data_ptr(), tr.S.data_ptr(),
                                tr.rowmax.data_ptr(), tr.rowsum.data_ptr(), tr.P.data_ptr(), None, tr.ws.data_ptr(),
                                tr.ws.numel(), s), "ep_fwd")
        E._lib.check(lib.ep_bwd_proj(tr.dout.data_ptr(), tr.P.data_ptr(), tr.out.data_ptr(), pool_mod.v.weight.data_ptr(), None, xt, B, N, D, M, 1,
                                     tr.g["v_w"].data_ptr(), None, tr.ws.data_ptr(), tr.ws.numel(), s), "ep_bwd_proj")
        E._lib.check(lib.ep_bwd_pool(x.data_ptr(), xt, pool_mod.cls_token.data_ptr(), float(pool_mod.scale), B, N, D, M,
                                     1, tr.S.data_ptr(), tr.rowmax.data_ptr(), tr.rowsum.data_ptr(),
                                     tr.g["cls"].data_ptr(), tr.ws.data_ptr(), tr.ws.numel(), s), "ep_bwd_pool")
        if i == 1:
            E._lib.kernel_timings()                                # drop the warm-up records
    fam = lib.ep_last_kernel_family()
    lib.ep_set_debug(args.debug_bits)
    agg = {}
    for nm, us in E._lib.kernel_timings():
        agg.setdefault(nm, []).append(us)
    kernels = {k: sum(v) / len(v) for k, v in agg.items()}
    streaming = {k: v for k, v in kernels.items() if k.startswith(("ks<", "kp<", "pool_", "fused "))}
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    tkey = f"{args.config}_M{M}"

    def roof(name, us):
        ach = alg_bytes / (us * 1e-6) / 1e9
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(tkey, {}).get(name), "kernel_us": us, "algorithmic_bytes": alg_bytes,
                "peak_source": peak_src}
    per_kernel = {k: roof(k, v) for k, v in streaming.items()}
    dominant = max(streaming, key=streaming.get) if streaming else None
    r_dom = per_kernel[dominant] if dominant else {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s",
                                                    "frac": None, "traffic": None}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(args, cfg, world),
            "run": {"comm_sms": tr.comm_sms, "cuda_graph": not args.no_graph, "kernel_family": fam,
                    "operand_copies_by_producers": tr.fuse_ops,
                    "l2": f"inputs larger than L2: {args.pool} resident batches x {alg_bytes / 1e6:.0f} MB cycled",
                    "replica_max_abs_diff": replica_diff},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(ms2) / e2e_steps,
                    "h2d_gbs_per_gpu": h2d / (float(ms2) / e2e_steps * 1e-3) / 1e9, "cpu_affinity": affinity},
            "gpu_launches": (tr.launches_per_step or 0) * args.steps,
            "launches_per_step": tr.launches_per_step,
            "roofline": r_dom, "roofline_all_streaming_kernels": per_kernel,
            "kernel_us": {k: round(v, 1) for k, v in kernels.items()},
            "step_roofline_frac": (2 * alg_bytes / (ms_per_step * 1e-3) / 1e9) / peak,
            "mean_loss": loss}
    if not args.no_cpu_baseline:
        try:
            sb = args.cpu_sample or 64
            val, sec, cores = cpu_reference_step(cfg, M, sb, iters=8, warmup=2)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sb} of {B} samples of the same workload per step, reference "
                                              f"formulation fwd+bwd+LARS, torch fp32, 8 timed steps, {sec * 1e3:.0f} ms/step"}
        except Exception as e:                                    # the baseline is a report, never a gate
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
