#!/usr/bin/env python
"""Benchmark of the SNVC dense stereo-to-voxel hot path on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the global-branch hot path (plane-sweep cost volume -> 3-D trunk
dres0/dres1/hourglass -> frustum-to-voxel lift) over one batch of 8 synthetic KITTI-shaped stereo
pairs (BASELINE.json configs[1]; features [8,32,96,312] fp32, 48 depth bins, voxel grid
[192,20,304]) per GPU.  Pairs are independent, so ranks shard them with no data-path collective
("scaling": "weak", 8 pairs per rank per step).  Prints ONE JSON line on rank 0.

`--impl reference` times the reference's CPU path for the same stages (multi-threaded torch ops,
oracle/torch_path.py; the cost-volume op has no CPU implementation in the reference so that stage
is the torch restatement) on the host cores, one pair per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

PAIRS_PER_GPU = 8
FEAT_C, FEAT_H, FEAT_W, DEPTH_BINS = 32, 96, 312, 48
TRUNK_GFLOP_PER_PAIR = 491.90          # BASELINE.md section 3 (sum 2*k^3*Cin*Cout*V_out)
CV_BYTES_PER_PAIR_BF16 = 191_692_992   # fp32 in -> bf16 NDHWC out (full 64-channel volume)
# split form (the product path): features in, right-half volume + three left planes out
CV_BYTES_PER_PAIR_SPLIT = 2 * FEAT_C * FEAT_H * FEAT_W * 4 + DEPTH_BINS * 4 + (DEPTH_BINS + 3) * FEAT_H * FEAT_W * FEAT_C * 2
CONV1_GFLOP_FULL = 2 * 27 * 64 * 32 * DEPTH_BINS * FEAT_H * FEAT_W * 1e-9      # dres0.conv1 on 64 channels, per pair (159.0)
CONV1_GFLOP_RIGHT = CONV1_GFLOP_FULL / 2                                        # its right-half launch (79.5)
ADDEND_GFLOP = 2 * 27 * 32 * 32 * 3 * FEAT_H * FEAT_W * 1e-9                   # the 3-plane addend convolution (5.0)
LIFT_VOX = 192 * 20 * 304


def lift_bytes_per_pair(out_bytes):
    # read the bf16 trunk output once (32*48*96*312*2) + write the lifted voxels
    return 32 * 48 * 96 * 312 * 2 + LIFT_VOX * 32 * out_bytes


def geometry():
    """CPU arm only: the oracle's own geometry object (independent of the product's)."""
    from oracle.global_branch import GlobalGeometry
    return GlobalGeometry()


def workload_config(world, eager=False):
    """`config` of the JSON line; both arms print the same one (the reference arm adds its sample)."""
    return {"workload": "global branch hot path, batch 8 synthetic KITTI pairs per GPU (feat 8x32x96x312 "
                        "fp32 -> cost volume 64x48x96x312 bf16 -> dres0/dres1/hourglass bf16 -> lift to "
                        "192x20x304 voxels), random init, BN eval folded",
            "pairs_per_gpu_per_step": PAIRS_PER_GPU, "parallelism": f"pair-sharded x{world}, no collective",
            "launch": "eager (one Python launch per kernel)" if eager else
                      "CUDA-graph replay (GraphedHotPath, one graph per stage, inputs copied device-to-device per step)",
            "cost_volume_form": "split (default): the depth-invariant left half of the 64-channel volume is kept as 3 planes "
                                "and enters dres0.conv1 as an addend; SNVC_SPLIT_CV=0 materialises the full volume",
            "l2": "inputs rotate over 4 sets (245 MB) and each step streams >3 GB of intermediates; "
                  "both exceed the 126 MB L2"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def cpu_reference_pairs_per_s(steps, warmup, threads=None):
    """The reference's CPU path (torch ops on the host cores), one pair per step."""
    import torch
    from oracle import blocks, torch_path
    import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    geom = geometry()
    trunk = blocks.GlobalTrunk(2 * FEAT_C, 32).eval()
    trunk.load_state_dict(synth.det_state_dict(trunk, 41))
    path = torch_path.GlobalHotPathCPU(trunk, geom)
    g = torch.Generator().manual_seed(10)
    left = torch.randn((1, FEAT_C, FEAT_H, FEAT_W), generator=g)
    right = torch.randn((1, FEAT_C, FEAT_H, FEAT_W), generator=g)
    shift = torch.from_numpy(geom.shifts(1))
    Ps = torch.from_numpy(geom.P[None].copy())
    for _ in range(warmup):
        path(left, right, shift, Ps)
    t0 = time.perf_counter()
    for _ in range(steps):
        path(left, right, shift, Ps)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps, threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    v, spp, threads = cpu_reference_pairs_per_s(steps, warmup)
    sample = f"{steps} timed + {warmup} warm-up passes of 1 synthetic pair (fp32, torch CPU ops)"
    line = {"impl": "reference", "metric": "stereo pairs/s (cost volume + 3D trunk + voxel lift)", "value": v,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": spp * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), sample="each step = 1 pair of that workload on the host cores"),
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def roofline_traffic():
    """DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import synth
    from snvc_b200 import _lib
    from snvc_b200.models.stereonet import GlobalHotPath, GraphedHotPath, HostPipeline
    from snvc_b200.extension.build_cost_volume import build_cost_volume_ndhwc_bf16
    from snvc_b200.utils.geometry import KITTI_P2, kitti_global_cfg, plane_sweep_shifts

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback in snvc_b200)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    cfg = kitti_global_cfg()
    model = GlobalHotPath(cfg).eval()
    model.load_state_dict(synth.det_state_dict(model, 41), strict=True)   # deterministic random init (tests/golden/synth.py)
    model = model.to(dev)
    B, K, W = PAIRS_PER_GPU, args.steps, args.warmup
    NSETS = 4   # inputs rotate over 4 sets (4 x 61 MB > 126 MB L2); intermediates (>3 GB / step) thrash L2 anyway
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    lefts = [torch.randn((B, FEAT_C, FEAT_H, FEAT_W), device=dev, generator=gen) for _ in range(NSETS)]
    rights = [torch.randn((B, FEAT_C, FEAT_H, FEAT_W), device=dev, generator=gen) for _ in range(NSETS)]
    shift = torch.from_numpy(plane_sweep_shifts(cfg, B)).to(dev)
    proj = torch.from_numpy(KITTI_P2[None].repeat(B, 0).copy()).to(dev)
    out_dtype, layout_out = torch.bfloat16, "NDHWC"
    os.environ["SNVC_B200_SKIP_SHIFT_CHECK"] = "1"

    ev = lambda: torch.cuda.Event(enable_timing=True)
    stage_ms = {"cost_volume": 0.0, "conv1": 0.0, "trunk": 0.0, "lift": 0.0}
    # The step is replayed from CUDA graphs (snvc_b200.models.stereonet.GraphedHotPath), one graph per stage -- cost volume |
    # 3-plane addend conv | dres0.conv1 | rest of the trunk | lift -- so that events between them time each stage inside
    # the timed region.
    # --eager launches the same kernels one by one from Python (host launch overhead then bounds the step).
    graphed = None if args.eager else GraphedHotPath(model, B, FEAT_C, (FEAT_H, FEAT_W), DEPTH_BINS, out_dtype,
                                                     layout_out, stages=True)

    def step(i, timed):
        l, r = lefts[i % NSETS], rights[i % NSETS]
        e = [ev() for _ in range(6)] if timed else None
        if graphed is not None:
            graphed.load(l, r, shift, proj)                 # device-to-device copy into the graph's input buffers
            if timed:
                e[0].record()
            # events: e0 start | e1 after the volume build | e5 after the addend conv (split form) | e2 after conv1 |
            # e3 after the rest of the trunk | e4 after the lift
            order = {"cost_volume": 1, "conv1_addend": 5, "conv1": 2, "trunk_rest": 3, "lift": 4}
            names = graphed.stage_names
            vox = graphed.replay((lambda k: e[order[names[k]]].record()) if timed else None)
            return vox, e
        if timed:
            e[0].record()
        cost = build_cost_volume_ndhwc_bf16(l, r, shift, 1)
        if timed:
            e[1].record()
        feat = model.trunk(cost, mark=(lambda name: e[2].record()) if timed else None)
        if timed:
            e[3].record()
        vox = model.lift(feat, proj, out_dtype, layout_out)
        if timed:
            e[4].record()
        return vox, e

    with torch.no_grad():
        for i in range(W):
            step(i, False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        t_start, t_stop = ev(), ev()
        evs = []
        launches0 = L.snvc_launch_count()
        t_start.record()
        for i in range(K):
            _, e = step(W + i, True)
            evs.append(e)
        t_stop.record()
        launches = L.snvc_launch_count() - launches0
        if graphed is not None:                             # replayed kernels do not pass the library's counter
            launches = K * graphed.launches_per_replay
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        elapsed_ms = t_start.elapsed_time(t_stop)
        has_addend_stage = graphed is not None and "conv1_addend" in graphed.stage_names
        for e in evs:
            stage_ms["cost_volume"] += e[0].elapsed_time(e[1])
            stage_ms["conv1"] += (e[5] if has_addend_stage else e[1]).elapsed_time(e[2])
            stage_ms["trunk"] += e[1].elapsed_time(e[3])        # (split form: includes the 3-plane addend convolution)
            stage_ms["lift"] += e[3].elapsed_time(e[4])

        # ---- end to end through the public host-buffer API: pinned host inputs -> H2D -> hot path -> D2H of
        # the lifted voxels into pinned host memory, every step, copies overlapped with compute on separate
        # streams (snvc_b200.models.stereonet.HostPipeline) -------------------------------------------------
        NH = 2
        h_in = [(torch.randn((B, FEAT_C, FEAT_H, FEAT_W)).pin_memory(), torch.randn((B, FEAT_C, FEAT_H, FEAT_W)).pin_memory(),
                 torch.from_numpy(plane_sweep_shifts(cfg, B)).pin_memory(),
                 torch.from_numpy(KITTI_P2[None].repeat(B, 0).copy()).pin_memory()) for _ in range(NH)]
        Z, Y, X = model.zs.numel(), model.ys.numel(), model.xs.numel()
        h_out = [torch.empty((B, Z, Y, X, 32), dtype=out_dtype).pin_memory() for _ in range(NH)]
        pipe = HostPipeline(model, depth=2, out_dtype=out_dtype, layout_out=layout_out, graphed=not args.eager)
        e2e_steps = max(4, min(K, 20))
        for i in range(3):
            pipe.submit(*h_in[i % NH], h_out[i % NH])
        pipe.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        cur = torch.cuda.current_stream()
        e0.record()
        for i in range(e2e_steps):
            pipe.submit(*h_in[i % NH], h_out[i % NH])
        cur.wait_stream(pipe.s_out)
        cur.wait_stream(pipe.s_in)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1)
        # same pipeline with the result left on the device (only a 64-value digest per pair returns to the host): what an
        # on-device consumer of the voxels sees; reported next to the dense-result number, never instead of it
        e2e_dev_ms = None
        if not args.eager:
            h_dig = [torch.empty((B, 64), dtype=out_dtype).pin_memory() for _ in range(NH)]
            for i in range(3):
                pipe.submit(*h_in[i % NH], h_dig[i % NH])
            pipe.drain()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e2, e3 = ev(), ev()
            e2.record()
            for i in range(e2e_steps):
                pipe.submit(*h_in[i % NH], h_dig[i % NH])
            cur.wait_stream(pipe.s_out)
            cur.wait_stream(pipe.s_in)
            e3.record()
            torch.cuda.synchronize()
            e2e_dev_ms = e2.elapsed_time(e3)
        clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([elapsed_ms, e2e_ms, stage_ms["cost_volume"], stage_ms["trunk"], stage_ms["lift"], stage_ms["conv1"],
                      e2e_dev_ms if e2e_dev_ms is not None else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, cv_ms, trunk_ms, lift_ms, conv1_ms, e2e_dev_ms = t.tolist()

    if rank == 0:
        peaks = measured_peaks()
        pairs = world * B * K
        value = pairs / (elapsed_ms * 1e-3)
        split = graphed is not None and graphed.split
        # executed FLOPs: the split first layer convolves the depth-constant left half once (3 planes) instead of 48 times
        trunk_gflop = TRUNK_GFLOP_PER_PAIR - (CONV1_GFLOP_FULL - CONV1_GFLOP_RIGHT) + ADDEND_GFLOP if split else TRUNK_GFLOP_PER_PAIR
        trunk_tflops = trunk_gflop * 1e-3 * B * K / (trunk_ms * 1e-3)
        conv1_gflop = CONV1_GFLOP_RIGHT if split else CONV1_GFLOP_FULL
        conv1_tflops = conv1_gflop * 1e-3 * B * K / (conv1_ms * 1e-3)
        cv_bytes = CV_BYTES_PER_PAIR_SPLIT if split else CV_BYTES_PER_PAIR_BF16
        cv_gbs = cv_bytes * B * K / (cv_ms * 1e-3) / 1e9
        lift_gbs = lift_bytes_per_pair(2) * B * K / (lift_ms * 1e-3) / 1e9
        traffic = roofline_traffic()
        cpu_v, cpu_spp, cpu_threads = cpu_reference_pairs_per_s(2, 1) if world == 1 and not args.no_cpu_baseline \
            else (None, None, None)
        line = {
            "metric": "stereo pairs/s (cost volume + 3D trunk + voxel lift)", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": elapsed_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, args.eager),
            "clocks": clocks,
            "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": int(2 * B * FEAT_C * FEAT_H * FEAT_W * 4 + B * DEPTH_BINS * 4 + B * 48),
                    "d2h_bytes_per_step": int(B * LIFT_VOX * 32 * 2), "steps": e2e_steps,
                    "result_on_device": ({"value": world * B * e2e_steps / (e2e_dev_ms * 1e-3), "unit": "pairs/s",
                                          "d2h_bytes_per_step": int(B * 64 * 2),
                                          "what": "same pipeline and H2D traffic; the voxels stay in HBM for an on-device "
                                                  "consumer, a 64-value digest per pair is read back"} if e2e_dev_ms else None),
                    "api": "snvc_b200.models.stereonet.HostPipeline.submit (pinned host buffers; H2D / compute / D2H "
                           "on three streams, 2 slots" + ("" if args.eager else ", one CUDA-graph replay per batch") + ")"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor",
                         "kernel": ("conv3d_kdpair_kernel<2,64,addend> (dres0.conv1 3x3x3 on the split cost volume: right half "
                                    "32->32 + depth-invariant addend, 1 launch / step)") if split else
                                   "conv3d_kdpair_kernel<4,128> (dres0.conv1 3x3x3 64->32, 1 launch / step)",
                         "achieved": conv1_tflops, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": conv1_tflops / peaks["tf_sustained"],
                         "traffic": traffic.get("dres0.conv1_split_dram_bytes_per_launch" if split else
                                                "dres0.conv1_dram_bytes_per_launch"),
                         "algorithmic_flop_per_launch": conv1_gflop * 1e9 * B,
                         "peak_source": peaks["source"] + " (sustained bf16, kernel timed inside a long step)",
                         "share_of_step": conv1_ms / elapsed_ms},
            "stages": {"cost_volume": {"ms_per_step": cv_ms / K, "achieved_gbs": cv_gbs, "frac_hbm": cv_gbs / peaks["hbm"],
                                       "bytes_per_pair": cv_bytes,
                                       "form": "split: right-half volume + 3 left planes (the 3-plane addend convolution is timed "
                                               "with the trunk)" if split else "full 64-channel volume"},
                       "trunk": {"ms_per_step": trunk_ms / K, "achieved_tflops": trunk_tflops,
                                 "frac_tensor": trunk_tflops / peaks["tf_sustained"], "share_of_step": trunk_ms / elapsed_ms,
                                 "executed_gflop_per_pair": trunk_gflop, "reference_gflop_per_pair": TRUNK_GFLOP_PER_PAIR},
                       "lift": {"ms_per_step": lift_ms / K, "achieved_gbs": lift_gbs, "frac_hbm": lift_gbs / peaks["hbm"]}},
        }
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": "pairs/s", "cores": cpu_threads, "kind": "port",
                                    "sample": "2 timed + 1 warm-up passes of 1 synthetic pair, same stages, fp32 torch "
                                              "CPU ops (oracle/torch_path.py)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch kernel by kernel from Python instead of CUDA-graph replay")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
