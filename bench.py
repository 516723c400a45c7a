#!/usr/bin/env python3
"""bench.py — ORB stereo frames/s of the B200 path (BASELINE.json metric), with roofline and CPU baseline.

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): a batch of KITTI-shaped stereo frames
(1241x376 left+right, ORBextractor(2000, 1.2, 8, 20, 7)), synthetic `orbx-synth-v1` texture, sharded by frame across the
GPUs with no collective (weak scaling: every rank processes --pairs stereo frames per step, default 4096).

  value : frames/s with the level-0 images already resident in HBM (7.8 GB per rank >> 126 MB L2), CUDA events on the
          launching stream, max over ranks.
  e2e   : the same metric through the public API with HOST buffers: pinned host images -> H2D -> kernels -> D2H of
          keypoints + descriptors + counts, three pipelined handles (streams); copies inside the timed region.
  --impl reference : the reference's CPU path (oracle/_ref when it was built, else the CPU oracle port) on all host cores.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1241, 376
ORB_ARGS = (2000, 1.2, 8, 20, 7)
METRIC = "ORB stereo frames/sec (KITTI 1241x376 left+right, 2000 kp/image, 8 levels, FAST 20/7)"
UNIT = "frames/s"
L2_BYTES = 126e6


WORKLOAD = "BASELINE.json configs[4]: batch of KITTI-shaped stereo frames 1241x376 (left+right), ORBextractor(2000,1.2,8,20,7), sharded by frame"


def level_sizes():
    import tc2li_slam_b200 as T
    return [(g["w"], g["h"]) for g in (T.level_geometry(W, H, l, *ORB_ARGS[:3]) for l in range(ORB_ARGS[2]))]


def algorithmic_bytes(kps_per_image):
    """compulsory bytes per IMAGE and per stage (each stage reads its input once, writes its output once)."""
    px = [w * h for w, h in level_sizes()]
    P, P0, P7 = sum(px), px[0], px[-1]
    K = kps_per_image
    return {"pyramid": (P - P7) + (P - P0), "fast": P, "blur": 2 * P, "octree": 0, "assemble": 60 * K, "orient_describe": (749 + 512) * K,
            "total": 5 * P - P0 - P7 + 1321 * K}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_backend():
    """oracle/_ref (verbatim reference TU compiled against stand-in OpenCV headers) when present, else the oracle port."""
    try:
        from oracle import pyref
        if pyref.available():
            return "reference", pyref.RefExtractor
    except Exception:
        pass
    from oracle import pyoracle
    return "port", pyoracle.OracleExtractor


def cpu_images(n_pairs, seed0):
    from oracle import pyoracle as po
    out = np.empty((2 * n_pairs, H, W), np.uint8)
    for f in range(n_pairs):
        out[2 * f] = po.synth_v1(W, H, seed0 + f)
        out[2 * f + 1] = po.synth_v1(W, H, seed0 + f, 16, 0x5bd1e995)
    return out


def cpu_run(images, cores, make_extractor, keep=None):
    """Extract every image, sharded over `cores` host threads (the C library releases the GIL).  Returns (seconds, keypoints);
    keep: a list of len(images) that receives (mono, keypoints, descriptors) per image (the parity check of the bench line)."""
    n = len(images)
    exs = [make_extractor(*ORB_ARGS) for _ in range(cores)]
    nk = [0] * cores

    def work(t):
        for i in range(t, n, cores):
            m, k, d = exs[t](images[i])
            nk[t] += len(k)
            if keep is not None:
                keep[i] = (m, k.copy(), d.copy())

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return time.perf_counter() - t0, sum(nk)


def workload_config(pairs):
    """`config` of both arms (identical dicts): the named workload, its size per GPU and step, and how the L2 is kept out of the timing."""
    return {"workload": WORKLOAD, "pairs_per_gpu_per_step": pairs,
            "l2": f"inputs {2 * pairs * W * H / 1e9:.2f} GB per rank and step, larger than the {L2_BYTES / 1e6:.0f} MB L2; no flush needed"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, make = cpu_backend()
    cores = os.cpu_count() or 1
    pairs = max(cores, min(args.ref_pairs, 64 * cores))
    imgs = cpu_images(pairs, 1000)
    for _ in range(args.warmup):
        cpu_run(imgs[: 2 * cores], cores, make)
    times = []
    for _ in range(args.steps):
        dt, _ = cpu_run(imgs, cores, make)
        times.append(dt)
    sec = sum(times) / len(times)
    v = pairs / sec
    sample = f"{pairs} synthetic KITTI-shaped stereo frames per step on {cores} host threads ({kind} of ORBextractor.cc, -O3, no FMA contraction)"
    emit_line(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.pairs),   # the same dict as the GPU arm's: what was actually timed of it is under cpu_baseline.sample
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        # under load = upper half of the samples (the sampler also sees the idle edges)
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """Everything any library writes to fd 1 from here on (NCCL prints its version there) goes to stderr; the one JSON
    line of the contract is written to the real stdout by emit_line()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit_line(text):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (text + "\n").encode())


def bind_to_gpu_numa_node(device):
    """One process per GPU: run on the cores of the NUMA node the GPU hangs off, so that the pinned staging buffers (first
    touch) and the copies that feed the GPU stay on its socket.  Returns a description, or None when the topology is not
    exposed (single-socket hosts, containers without sysfs)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(device), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        return None


def feed_ceiling(devices, h2d_bytes, d2h_bytes, frames):
    """The host <-> device copy ceiling for this step's traffic on these GPUs: profiles/tools/feed_probe run live (pinned buffers of
    the benchmark's chunk sizes, one host process per GPU), else the committed measurement profiles/r02/feed_ceiling.json."""
    run, src = None, None
    exe = os.path.join(ROOT, "profiles", "tools", "feed_probe")
    try:
        if os.access(exe, os.X_OK):
            out = subprocess.run([exe, "0.3", ",".join(str(d) for d in devices)], capture_output=True, text=True, timeout=120).stdout
            run, src = json.loads(out)["runs"][0], "live: profiles/tools/feed_probe (one process per GPU)"
            if not run.get("ok", True):
                run = None
    except Exception:
        run = None
    if run is None:
        try:
            for r_ in json.load(open(os.path.join(ROOT, "profiles", "r02", "feed_ceiling.json")))["runs"]:
                if r_["gpus"] == list(devices):
                    run, src = r_, "profiles/r02/feed_ceiling.json"
        except Exception:
            pass
    if run is None:
        return None
    # Two models.  The CEILING is the upper one: every byte of the step's input crosses at the H2D-only rate with the (four times smaller) output fully hidden
    # behind it — no implementation can beat it on this box.  The duplex figure is what the link gives when BOTH directions are kept saturated all the time (the
    # probe rate-matches them); the real pipeline sends 1 byte back per 3.8 bytes in, so it lands between the two (at eight GPUs it exceeds the duplex model).
    t_hi = max(h2d_bytes / (run["h2d_only_GBps"] * 1e9), d2h_bytes / (run["d2h_only_GBps"] * 1e9))
    t_lo = max(h2d_bytes / (run["duplex_h2d_GBps"] * 1e9), d2h_bytes / (run["d2h_only_GBps"] * 1e9))
    return {"source": src, "duplex_h2d_GBps": run["duplex_h2d_GBps"], "d2h_only_GBps": run["d2h_only_GBps"], "h2d_only_GBps": run["h2d_only_GBps"],
            "ceiling_frames_per_s": frames / t_hi, "duplex_model_frames_per_s": frames / t_lo,
            "what": "ceiling: frames/s if the step's input crossed at the measured H2D-only rate and nothing else took time; duplex model: the same with both "
                    "directions saturated all the time"}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=4096, help="stereo frames per rank per step")
    ap.add_argument("--group", type=int, default=0, help="images per launch set (chunk) of the resident run, chunks round-robin over 3 streams (0 = all images in one launch set: measured fastest)")
    ap.add_argument("--chunk", type=int, default=128, help="images per pipelined handle of the e2e run (measured: 256 -> 55.8k, 128 -> 57.0k, 64 -> 49.2k frames/s)")
    ap.add_argument("--ref-pairs", type=int, default=256, help="stereo frames per step of the CPU arms (bounded sample)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-stereo", action="store_true")
    ap.add_argument("--no-large", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing hygiene: at least three warm-up steps

    if args.impl == "reference":
        reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # rank -> GPU: spread over the box (N = 2 -> GPUs 0, 4; N = 4 -> 0, 2, 4, 6 on an 8-GPU box).  Neighbouring GPUs share an upstream PCIe
    # port here: GPUs 0-3 together get what GPUs 0 and 4 get as a pair (profiles/r02/feed_ceiling.json)
    ndev = 0
    try:
        ndev = len([l_ for l_ in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout.splitlines() if l_.startswith("GPU ")])
    except Exception:
        pass
    spread = ndev // world if ndev >= world and world > 1 and ndev % world == 0 else 1
    devices = [r_ * spread for r_ in range(world)]
    local = devices[local] if world > 1 else local
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        # the data path has no collective (frames are independent): the ranks only meet at start barriers and for the max / sum of
        # a few floats, on the host (gloo); NCCL is not initialised
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo")

    def barrier():
        if dist is not None:
            dist.barrier()

    def _reduce(v, op):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(v):
        return _reduce(v, dist.ReduceOp.MAX) if dist is not None else v

    def sum_over_ranks(v):
        return _reduce(v, dist.ReduceOp.SUM) if dist is not None else v

    import tc2li_slam_b200 as T
    from tc2li_slam_b200.orbx import Event

    from tc2li_slam_b200.sharding import frame_seed, weak_scaling_frames
    n_img = 2 * args.pairs
    seed0 = frame_seed(weak_scaling_frames(args.pairs, rank)[0])  # every rank extracts its own, distinct frames
    group = min(args.group, n_img) if args.group > 0 else n_img

    # ---- resident run: all level-0 images live in the handle's pyramid buffer ----
    big = T.BatchExtractor(W, H, n_img, *ORB_ARGS, device=local)   # (local = this rank's GPU)
    big.synth(seed0, n_img, stereo=True)
    big.sync()

    def resident_step():
        big.run_pipelined(n_img, group)  # chunks of `group` images round-robin over the handle's three internal streams

    for _ in range(args.warmup):
        resident_step()
    big.check()
    counts, _ = big.download_counts(n_img)
    kp_mean = float(counts.mean())
    ev0, ev1 = Event(big), Event(big)
    big.sync()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0.record(big)
    for _ in range(args.steps):
        resident_step()
    ev1.record(big)
    ms_total = ev0.elapsed_ms(ev1)   # (waits for the stop event: the device is idle again)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_step = max_over_ranks(ms_total / args.steps)
    value = args.pairs * world / (ms_step * 1e-3)

    # ---- per-stage device time of one step (events between the stages), for the roofline ----
    big.enable_timing(True)
    stage_ms = {k: 0.0 for k in T.orbx.STAGE_NAMES}
    reps = 2
    for _ in range(reps):
        for first in range(0, n_img, group):
            big.run_range(first, min(group, n_img - first))
            for k, v in big.stage_ms().items():
                stage_ms[k] += v / reps
    big.enable_timing(False)
    launches_per_step = big.kernel_launches * ((n_img + group - 1) // group)
    ab = algorithmic_bytes(kp_mean)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    stages = {}
    for k, ms in stage_ms.items():
        gbs = ab[k] * n_img / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stages[k] = {"ms": ms, "bytes_per_image": ab[k], "GBps": gbs, "frac": gbs / peak}
    dom = max(stage_ms, key=stage_ms.get)
    launches_of = {"pyramid": ORB_ARGS[2], "fast": 1, "blur": ORB_ARGS[2], "octree": 2, "assemble": 1, "orient_describe": 1}
    n_sets = (n_img + group - 1) // group
    dom_launch_ms = stage_ms[dom] / (launches_of[dom] * n_sets)
    dom_bytes_launch = ab[dom] * n_img / (launches_of[dom] * n_sets)
    traffic = None  # measured DRAM read+write bytes per launch of the dominant kernel (ncu --set full, profiles/traffic.json)
    issue = None    # the bound the dominant kernel actually runs against when it is not HBM: instruction issue (4 warp instructions per clock and SM)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj["bytes_per_image"][dom] * n_img / (launches_of[dom] * n_sets)
        wi = tj["warp_instructions_per_image"][dom]
        px = sum(w_ * h_ for w_, h_ in level_sizes())
        sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        issue = {"lane_instr_per_pixel": wi * 32.0 / px, "warp_instr_per_image": wi, "issue_pct": tj["issue_pct"][dom],
                 "achieved_occupancy_pct": tj["achieved_occupancy_pct"][dom],
                 "frac_of_issue_peak": wi * n_img / (stage_ms[dom] * 1e-3 * 148 * 4 * sm_mhz * 1e6),
                 "what": "warp instructions (smsp__inst_executed, ncu capture: " + tj["source"] + ") per image x images of this run / (stage time measured in this run "
                         "x 148 SMs x 4 schedulers x the SM clock sampled in this run); issue_pct is ncu's own smsp__issue_active of that capture"}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": dom_bytes_launch / (dom_launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": dom_bytes_launch / (dom_launch_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "issue": issue, "peak_source": peak_src,
                "bytes_per_launch": dom_bytes_launch, "ms_per_launch": dom_launch_ms,
                "whole_step": {"bytes_per_frame": 2 * ab["total"], "GBps": 2 * ab["total"] * args.pairs / (ms_step * 1e-3) / 1e9,
                               "frac": 2 * ab["total"] * args.pairs / (ms_step * 1e-3) / 1e9 / peak}}

    # ---- supplementary: the next step of the caller, Frame::ComputeStereoMatches, on the resident results (not part of `value`) ----
    stereo = None
    if not args.no_stereo:
        mb, mbf = 0.5372, 387.5744  # KITTI 00-02: baseline 0.537 m, bf 387.57 (config/Camera-Lidar/KITTI00-02.yaml)
        big.stereo_match(args.pairs, mb, mbf)
        big.sync()
        s0, s1 = Event(big), Event(big)
        s0.record(big)
        for _ in range(args.steps):
            big.stereo_match(args.pairs, mb, mbf)
        s1.record(big)
        ms_st = s0.elapsed_ms(s1) / args.steps
        u, _ = big.download_stereo(min(args.pairs, 64))
        stereo = {"what": "orbx_batch_stereo_match (Frame::ComputeStereoMatches, Frame.cc:841-1011) on the resident keypoints/descriptors/pyramids",
                  "ms_per_step": ms_st, "frames_per_s": args.pairs / (ms_st * 1e-3), "matched_fraction_first_64": float((u >= 0).sum()) / float(max(1, (counts[:2 * min(args.pairs, 64):2]).sum()))}

    # ---- supplementary: the feature grid every Frame constructor builds next (Frame::AssignFeaturesToGrid) + area queries ----
    grid = None
    if not args.no_stereo:
        big.assign_grid(n_img)
        big.sync()
        s0, s1 = Event(big), Event(big)
        s0.record(big)
        for _ in range(args.steps):
            big.assign_grid(n_img)
        s1.record(big)
        ms_grid = s0.elapsed_ms(s1) / args.steps
        nq = 1 << 16  # one 15-px window per query, spread over the first 64 images (a projection search issues ~1-2 k per frame)
        rng = np.random.default_rng(11)
        q = np.zeros(nq, T.AREA_QUERY_DTYPE)
        q["image"] = rng.integers(0, min(n_img, 64), nq)
        q["x"], q["y"], q["r"] = rng.uniform(0, W, nq), rng.uniform(0, H, nq), 15.0
        q["min_level"], q["max_level"] = -1, -1
        big.features_in_area(q, 64)
        t0 = time.perf_counter()
        qc, _ = big.features_in_area(q, 64)
        dt_q = time.perf_counter() - t0
        # projection search of one frame: the keypoints of image 0 re-projected onto themselves (th = 15, default level window)
        kp0, de0, cn0, _ = big.download(1)
        n0 = int(cn0[0])
        sfs = T.tables(2000, 1.2, 8, 20, 7)["scale_factors"]
        pq = np.zeros(n0, T.PROJ_QUERY_DTYPE)
        pq["u"], pq["v"], pq["angle"] = kp0[0, :n0]["x"], kp0[0, :n0]["y"], kp0[0, :n0]["angle"]
        pq["radius"] = np.float32(15.0) * sfs[kp0[0, :n0]["octave"]]
        pq["ur"], pq["blocks"] = -1.0, 1
        pq["min_level"], pq["max_level"] = kp0[0, :n0]["octave"] - 1, kp0[0, :n0]["octave"] + 1
        pq["desc"] = de0[0, :n0]
        big.search_by_projection(0, pq, False, True)
        t0 = time.perf_counter()
        for _ in range(10):
            _, nm_sp = big.search_by_projection(0, pq, False, True)
        dt_sp = (time.perf_counter() - t0) / 10
        grid = {"projection_search": {"what": "orbx_batch_search_by_projection (ORBmatcher::SearchByProjection from ORBmatcher.cc:1729 on): the keypoints of one frame "
                                              "re-projected onto themselves, th = 15, host queries in, host matches out", "map_points": n0, "matches": int(nm_sp),
                                      "ms_per_frame_host_to_host": dt_sp * 1e3},
                "what": "orbx_batch_assign_grid (Frame::AssignFeaturesToGrid, Frame.cc:412-443) on the resident keypoints; orbx_batch_features_in_area "
                        "(Frame::GetFeaturesInArea, Frame.cc:687-749) host queries in, host index lists out",
                "assign_ms_per_step": ms_grid, "assign_images_per_s": n_img / (ms_grid * 1e-3), "area_queries": nq,
                "area_queries_per_s_host_to_host": nq / dt_q, "mean_hits_per_query": float(qc.mean())}

    # ---- supplementary: Frame::ComputeBoW on the resident descriptors (DBoW2 transform; synthetic vocabulary of ORBvoc.txt's shape: k = 10, L = 6) ----
    bow = None
    if not args.no_stereo:
        k_voc, L_voc = 10, 6
        n_nodes = (k_voc ** (L_voc + 1) - 1) // (k_voc - 1)
        rng = np.random.default_rng(5)
        parent = np.zeros(n_nodes, np.int32)
        parent[1:] = (np.arange(1, n_nodes) - 1) // k_voc          # complete k-ary tree in breadth-first numbering
        first_leaf = n_nodes - k_voc ** L_voc
        is_word = (np.arange(n_nodes) >= first_leaf).astype(np.uint8)
        vdesc = rng.integers(0, 256, (n_nodes, 32), dtype=np.uint8)
        weight = np.where(is_word == 1, rng.uniform(0.5, 12.0, n_nodes), 0.0)
        voc = T.ORBVocabulary(local).from_arrays(k_voc, L_voc, 0, 0, parent, is_word, vdesc, weight)
        big.compute_bow(voc, n_img, 4)
        big.sync()
        s0, s1 = Event(big), Event(big)
        s0.record(big)
        for _ in range(args.steps):
            big.compute_bow(voc, n_img, 4)
        s1.record(big)
        ms_bow = s0.elapsed_ms(s1) / args.steps
        r0 = big.download_bow(1)[0]
        bow = {"what": "orbx_batch_compute_bow (Frame::ComputeBoW, Frame.cc:768-775 = DBoW2 TemplatedVocabulary::transform, levelsup 4) on the resident descriptors; "
                       "synthetic vocabulary k = 10, L = 6 (10^6 words, 35.6 MB of node descriptors: L2 resident); the real ORBvoc.txt is absent from the reference checkout",
               "ms_per_step": ms_bow, "images_per_s": n_img / (ms_bow * 1e-3), "descriptors_per_s": float(counts.sum()) / (ms_bow * 1e-3),
               "words_first_image": int(len(r0[0])), "nodes_first_image": int(len(r0[2]))}
        voc.close()

    # ---- e2e: the product call.  Pinned host images -> orbx_pool_extract -> host keypoints / descriptors / counts, frame order.
    #      N = 1: a pool of this GPU.  N > 1: every rank first drives its own GPU at the same time (supplementary), then rank 0
    #      alone drives ALL N GPUs through ONE pool from one process — the multi-GPU product API — while the other ranks wait ----
    e2e = None
    e2e_one = None
    pinned_in = None
    out_kps = out_desc = out_cnt = out_mono = None
    feed = None
    if not args.no_e2e:
        chunk = min(args.chunk, n_img)
        K = big.K
        pinned_in = T.PinnedArray((n_img, H, W), np.uint8)
        for first in range(0, n_img, 1024):
            n = min(1024, n_img - first)
            big.download_images(first, n, pinned_in.array[first:first + n])
        out_kps = T.PinnedArray((n_img, K), T.KP_DTYPE)
        out_desc = T.PinnedArray((n_img, K, 32), np.uint8)
        out_cnt = T.PinnedArray((n_img,), np.int32)
        out_mono = T.PinnedArray((n_img,), np.int32)
        h2d = n_img * H * W
        d2h = n_img * (K * 28 + K * 32 + 8)

        def pool_run(pool, src, dst):
            for _ in range(args.warmup):
                pool.extract(src, None, *dst)
            ms = 0.0
            for _ in range(args.steps):
                pool.extract(src, None, *dst)
                ms += float(pool.last_ms().max())   # device time of the call: first upload enqueued -> last download done, slowest GPU
            return ms / args.steps

        pool = T.ExtractorPool(W, H, [local], *ORB_ARGS, chunk_images=chunk, handles_per_device=3)
        barrier()
        ms_mine = pool_run(pool, pinned_in.array, (out_kps.array, out_desc.array, out_cnt.array, out_mono.array))
        assert np.array_equal(out_cnt.array, counts), "e2e keypoint counts differ from the resident run"
        launches_e2e = pool.kernel_launches
        pool.close()
        barrier()
        ms_ranks = max_over_ranks(ms_mine)
        e2e = {"value": args.pairs * world / (ms_ranks * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": ms_ranks, "api": "orbx_pool_extract (include/orbx.h), one rank per GPU" if world > 1 else "orbx_pool_extract (include/orbx.h)",
               "pipeline": f"per GPU: one host thread, 3 handles x {chunk} images, pinned host buffers", "gpu_launches": launches_e2e * world,
               "gpus": devices if world > 1 else [local]}
        if world > 1:
            if rank == 0:
                feed = feed_ceiling(devices, h2d * world, d2h * world, args.pairs * world)   # (the other ranks wait at the barrier below)
                # supplementary: ONE process driving all N GPUs through one pool; GPU d's share are rank 0's frames again
                # (throughput does not depend on the texture)
                all_in = T.PinnedArray((world * n_img, H, W), np.uint8)
                for r_ in range(world):
                    all_in.array[r_ * n_img:(r_ + 1) * n_img] = pinned_in.array
                all_out = (T.PinnedArray((world * n_img, K), T.KP_DTYPE), T.PinnedArray((world * n_img, K, 32), np.uint8),
                           T.PinnedArray((world * n_img,), np.int32), T.PinnedArray((world * n_img,), np.int32))
                pool = T.ExtractorPool(W, H, devices, *ORB_ARGS, chunk_images=chunk, handles_per_device=3)
                ms_pool = pool_run(pool, all_in.array, tuple(a.array for a in all_out))
                assert np.array_equal(all_out[2].array, np.tile(counts, world)), "pool keypoint counts differ from the resident run"
                e2e_one = {"value": args.pairs * world / (ms_pool * 1e-3), "unit": UNIT, "ms_per_step": ms_pool,
                           "what": f"orbx_pool_extract from ONE process over all {world} GPUs (chunk c -> GPU c mod {world}); the driver serialises part of the "
                                   "copy submission inside a process: see profiles/r02/feed_ceiling_one_process.json against feed_ceiling.json",
                           "per_gpu_ms": [float(x) for x in pool.last_ms()]}
                pool.close()
                for a in (all_in,) + all_out:
                    a.free()
            barrier()
        elif rank == 0:
            feed = feed_ceiling([local], h2d, d2h, args.pairs)
        if feed is not None and feed.get("ceiling_frames_per_s"):
            e2e["frac_of_feed_ceiling"] = e2e["value"] / feed["ceiling_frames_per_s"]
            e2e["feed_ceiling"] = feed

    # ---- supplementary: BASELINE.json configs[3], large single frames (1920x1080 and 3840x2160, 4000 features) ----
    large = None
    if rank == 0 and not args.no_large:
        large = {}
        for (lw, lh, nimg) in ((1920, 1080, 64), (3840, 2160, 16)):
            lb = T.BatchExtractor(lw, lh, nimg, 4000, 1.2, 8, 20, 7, device=local)
            lb.synth(7, nimg, stereo=False)
            for _ in range(3):
                lb.run(nimg)
            lb.check()
            cl, _ = lb.download_counts(nimg)
            l0, l1 = Event(lb), Event(lb)
            l0.record(lb)
            for _ in range(args.steps):
                lb.run(nimg)
            l1.record(lb)
            ms = l0.elapsed_ms(l1) / args.steps
            px = [g_["w"] * g_["h"] for g_ in (T.level_geometry(lw, lh, l, 4000, 1.2, 8) for l in range(8))]
            P_, K_ = sum(px), float(cl.mean())
            bytes_img = 5 * P_ - px[0] - px[-1] + 1321 * K_
            large[f"{lw}x{lh}"] = {"images_per_launch_set": nimg, "ms_per_image": ms / nimg, "images_per_s": nimg / (ms * 1e-3), "keypoints_per_image": K_,
                                   "algorithmic_bytes_per_image": bytes_img, "GBps": bytes_img * nimg / (ms * 1e-3) / 1e9,
                                   "frac_of_measured_hbm": bytes_img * nimg / (ms * 1e-3) / 1e9 / peak}
            lb.close()

    # ---- supplementary: BASELINE.json configs[2], EuRoC-shaped 752x480 stereo with lapping windows, 1200 features, plus the
    #      descriptor stage of Frame::ComputeStereoFishEyeMatches (kNN(2) + Lowe gate over the lapping tails) ----
    euroc = None
    if rank == 0 and not args.no_large:
        ef = 1024
        eb = T.BatchExtractor(752, 480, 2 * ef, 1200, 1.2, 8, 20, 7, device=local)
        eb.synth(11, 2 * ef, stereo=True)
        lap = np.tile(np.array([[250, 751], [0, 501]], np.int32), (ef, 1))
        for _ in range(3):
            eb.run(2 * ef, lap)
            eb.lapping_match(ef)
        eb.check()
        e0, e1, e2_ = Event(eb), Event(eb), Event(eb)
        e0.record(eb)
        for _ in range(args.steps):
            eb.run(2 * ef, lap)
        e1.record(eb)
        for _ in range(args.steps):
            eb.lapping_match(ef)
        e2_.record(eb)
        ce, me = eb.download_counts(2 * ef)
        knn, okf = eb.download_lapping(8)
        nq = (ce[0:16:2] - me[0:16:2])
        euroc = {"workload": "1024 EuRoC-shaped stereo frames 752x480, ORBextractor(1200,1.2,8,20,7), lapping windows {250,751} / {0,501}",
                 "extract_frames_per_s": ef / (e0.elapsed_ms(e1) / args.steps * 1e-3), "keypoints_per_image": float(ce.mean()),
                 "lapping_queries_per_frame": float((ce[0::2] - me[0::2]).mean()),
                 "lapping_match_ms_per_1024_frames": e1.elapsed_ms(e2_) / args.steps,
                 "ratio_pass_fraction_first_8": float(sum(int(okf[f, :nq[f]].sum()) for f in range(8))) / float(max(1, int(nq.sum())))}
        eb.close()

    # ---- stereo-pair latency through the reference-shaped API (BASELINE.json configs[1]) ----
    latency = None
    if rank == 0 and not args.no_latency:
        st = T.StereoORBextractor(*ORB_ARGS, device=local)
        if pinned_in is not None:
            left, right = pinned_in.array[0], pinned_in.array[1]
        else:
            tmp = np.empty((2, H, W), np.uint8)
            big.download_images(0, 2, tmp)
            left, right = tmp[0], tmp[1]
        for _ in range(20):
            st(left, right)
        lat = []
        for _ in range(200):
            t0 = time.perf_counter()
            st(left, right)
            lat.append(time.perf_counter() - t0)
        lat = np.sort(np.array(lat)) * 1e3
        latency = {"stereo_pair_ms": float(np.median(lat)), "p10_ms": float(lat[20]), "p90_ms": float(lat[180]), "calls": 200,
                   "what": "orbx_extract_stereo, host images in, host keypoints+descriptors out, 2 streams (median of 200 calls through the Python mirror)"}
        st.close()

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only) ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        kind, make = cpu_backend()
        cores = os.cpu_count() or 1
        if pinned_in is not None:
            src = pinned_in.array
        else:
            src = np.empty((min(n_img, 64), H, W), np.uint8)
            big.download_images(0, len(src), src)
        dt1, _ = cpu_run(src[:2], 1, make)
        per_img = dt1 / 2
        n_s = int(min(len(src), max(2 * cores, (15.0 * cores / per_img) // 2 * 2)))
        kept = [None] * n_s
        dt, nk = cpu_run(src[:n_s], cores, make, kept)
        cpu = {"value": (n_s / 2) / dt, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"first {n_s // 2} stereo frames of this workload, sharded over {cores} host threads, {dt:.1f} s wall; pixel primitives are the "
                         f"oracle's scalar restatements of OpenCV's (a SIMD OpenCV build would be faster per core)",
               "single_thread_ms_per_image": per_img * 1e3}
        if out_kps is not None:
            # parity at benchmark scale: the CPU leg's keypoints / descriptors against what the timed e2e call left in the pinned outputs
            bad = 0
            for i_, (m_, k_, d_) in enumerate(kept):
                c_ = int(out_cnt.array[i_])
                if c_ != len(k_) or int(out_mono.array[i_]) != m_ or out_kps.array[i_, :c_].tobytes() != k_.tobytes() or not np.array_equal(out_desc.array[i_, :c_], d_):
                    bad += 1
            parity = {"frames_checked": n_s, "mismatches": bad, "against": f"CPU {kind} of ORBextractor.cc, every field of every keypoint and every descriptor byte"}

    if world > 1 and out_kps is not None:
        # every rank checks four of its own frames against the CPU path (outside every timed region)
        kind, make = cpu_backend()
        ex_ = make(*ORB_ARGS)
        bad = 0
        picks = [0, 1, n_img // 2, n_img - 1]
        for i_ in picks:
            m_, k_, d_ = ex_(pinned_in.array[i_])
            c_ = int(out_cnt.array[i_])
            if c_ != len(k_) or int(out_mono.array[i_]) != m_ or out_kps.array[i_, :c_].tobytes() != k_.tobytes() or not np.array_equal(out_desc.array[i_, :c_], d_):
                bad += 1
        parity = {"frames_checked": int(sum_over_ranks(len(picks))), "mismatches": int(sum_over_ranks(bad)),
                  "against": f"CPU {kind} of ORBextractor.cc on four frames of every rank"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.pairs),
            "run": {"images_per_launch_set": group, "keypoints_per_image": kp_mean},
            "clocks": clocks, "numa_binding": numa, "e2e": e2e, "e2e_one_process_all_gpus": e2e_one, "parity": parity, "gpu_launches": launches_per_step, "roofline": roofline, "cpu_baseline": cpu, "stages": stages, "latency": latency, "stereo_match": stereo, "feature_grid": grid, "bag_of_words": bow, "large_images": large, "euroc_lapping": euroc,
        }
        emit_line(json.dumps(line))
    big.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
