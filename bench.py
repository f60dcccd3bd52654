#!/usr/bin/env python
"""bench.py — clonotype-join throughput on B200 (BASELINE.json metric: exact-subclonotype pairs scored/sec; join wall time at
1M / 10M cells on 1/2/4/8 GPUs).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (bucket scan -> bucketer -> pair kernels -> forest -> union-find) over the whole synthetic
repertoire.  Workload at EVERY N: BASELINE configs[4], the 10M-cell BCR repertoire ("bucket-sharded across 1/2/4/8 B200"), i.e.
STRONG scaling: the same 4.5e10 pairs at every N.  At N > 1 every rank (one process per GPU) holds the whole input, joins the
sub-buckets / row ranges the library assigns to it (ejb_set_shard: deterministic LPT, identical on every rank, no collective on
the data path), the partial forests are gathered GPU-to-GPU over NCCL into rank 0, merged and finalized there on the device.
The N = 1 line additionally carries `configs1`: the 1M-cell configs[1] measured the same way in the same job (the workload the
per-kernel roofline work is tracked on), with the WHOLE-workload CPU baseline next to it.

  value     pairs/s with inputs resident in HBM (ejb_run [+ gather + merge]), device time
  e2e       pairs/s through the public API from HOST (pinned) buffers: H2D + partition + run + gather/merge + D2H + EquivRel replay
  roofline  popc pipe: algorithmic popc / time vs the popc peak measured in the same job, plus the EXECUTED popc of stage 1
  cpu_baseline  the CPU oracle (port of the reference join) on the WHOLE configs[1] workload, all host cores

--impl reference times the reference's own CPU algorithm (the oracle port; the Rust reference cannot be built in this image) on
a bounded, seeded bucket sample of the same workload per step, and prints the measured bias of that sampling rule.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "exact-subclonotype pairs scored/sec"
MAIN_CONFIG, SIDE_CONFIG = 4, 1
WORKLOADS = {
    4: "BASELINE configs[4]: synthetic 10M-cell BCR repertoire (configs[1] clone-size law, 16 datasets, 4 donors)",
    1: "BASELINE configs[1]: synthetic 1M-cell human BCR repertoire, heavy-tailed clone sizes (Zipf 1.4, <=50k cells/clone)",
}


def bucket_table(inp):
    """Per-row bucket id (runs of equal lens, join.rs:68-88)."""
    a = inp.a
    n = inp.n_rows
    lens = a["row_len"].reshape(-1, 2)
    nch = a["row_nchains"]
    head = np.ones(n, dtype=bool)
    head[1:] = (lens[1:] != lens[:-1]).any(axis=1) | (nch[1:] != nch[:-1])
    return np.cumsum(head) - 1, lens, nch


def algorithmic_popc(inp):
    """P1 = ceil(c_h/32)+ceil(c_l/32) popc per pair (stage 1);
    P2 = 3*(ceil(L_h/32)+ceil(L_l/32)) + 6 per stage-2 candidate.  Pair-weighted means over sub-buckets."""
    bucket, lens, nch = bucket_table(inp)
    c3 = inp.a["row_cdr3_len"].reshape(-1, 2)
    act = nch == 2
    key = (bucket[act].astype(np.int64) << 44) | (c3[act, 0].astype(np.int64) << 33) | (c3[act, 1].astype(np.int64) << 22) | \
          (lens[act, 0].astype(np.int64) << 11) | lens[act, 1].astype(np.int64)
    uniq, cnt = np.unique(key, return_counts=True)
    ch, cl, lh, ll = (uniq >> 33) & 0x7ff, (uniq >> 22) & 0x7ff, (uniq >> 11) & 0x7ff, uniq & 0x7ff
    pairs = cnt.astype(np.float64) * (cnt - 1) / 2
    p1 = np.ceil(ch / 32) + np.ceil(cl / 32)
    p2 = 3 * (np.ceil(lh / 32) + np.ceil(ll / 32)) + 6
    tot = max(pairs.sum(), 1.0)
    row_bytes = ((np.ceil(ch / 4) + np.ceil(cl / 4)) + 3 * 16 * (np.ceil(lh / 128) + np.ceil(ll / 128)) + 112) * cnt   # CDR3 planes + 3-plane records + aux
    return float((p1 * pairs).sum() / tot), float((p2 * pairs).sum() / tot), float(pairs.sum()), float(row_bytes.sum())


def cpu_sample(inp, row_fraction, seed=0, cap_rows=None):
    """A bounded sample of the workload for the per-step reference arm: lens-buckets in a seeded shuffle — NO bucket is excluded
    for its size — until they hold `row_fraction` of the rows.  With `cap_rows` a bucket contributes at most its first
    `cap_rows` rows (the reference runs one thread per bucket, so one 30k-row bucket alone is 18 s per pass: a 25-pass arm can
    only afford a prefix of it; the prefix keeps the bucket's clone structure, and `sample_bias` measures what the rule costs)."""
    bucket, lens, nch = bucket_table(inp)
    nb = int(bucket[-1]) + 1
    sizes = np.bincount(bucket, minlength=nb).astype(np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])   # rows of a lens-bucket are contiguous (info is sorted by lens)
    two = np.zeros(nb, dtype=bool)
    two[bucket[nch == 2]] = True
    target = row_fraction * inp.n_rows
    chosen, acc, capped = [], 0, 0
    for b in np.random.default_rng(seed).permutation(nb):
        if not two[b] or sizes[b] < 2:
            continue
        take = int(sizes[b]) if cap_rows is None else int(min(sizes[b], cap_rows))
        capped += take < sizes[b]
        chosen.append((int(b), take))
        acc += take
        if acc >= target:
            break
    chosen.sort()
    rows = np.concatenate([np.arange(starts[b], starts[b] + take) for b, take in chosen]) if chosen else np.zeros(0, dtype=np.int64)
    desc = (f"{len(chosen)} of {nb} lens-buckets (seeded shuffle, none excluded; largest {max((t for _, t in chosen), default=0)} rows"
            + (f"; {capped} truncated to their first {cap_rows} rows" if cap_rows is not None else "") + f"), {len(rows)} of {inp.n_rows} rows")
    return inp.take_rows(rows), desc


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def spread(xs):
    xs = sorted(xs)
    n = len(xs)
    return {"min": xs[0], "median": statistics.median(xs), "p95": xs[min(n - 1, int(0.95 * n))], "max": xs[-1], "n": n}


# --------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle (port of the Rust join) on a bounded sample per step
# --------------------------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank):
    if rank != 0:
        return
    import enclone_ranger_b200 as E
    from oracle import pyoracle
    p = E.default_params(True)
    empty = E.JoinInput()
    pyoracle.run(p, empty.c_ptr(), threads=0)  # builds the Stirling table outside the timed region (start.rs:344 does too)
    # sampling bias, measured where the whole workload is affordable: configs[1] in full vs the same sampling rule on it
    rep = E.generate(SIDE_CONFIG)
    side = rep.to_join_input()
    rep.close()
    full = pyoracle.run(p, side.c_ptr(), threads=0)
    rep = E.generate(MAIN_CONFIG, n_cells=args.cells)
    inp = rep.to_join_input()
    rep.close()
    # The sample is sized so that the W + K passes together take about a minute.  The reference's parallelism is one thread
    # per lens-bucket (join.rs:582), so a pass lasts as long as its slowest bucket: a bucket contributes at most its first
    # `cap` rows, and about one bucket per host thread is taken.  One calibration pass, then cap is rescaled with the square
    # root of the time ratio (pairs grow with rows squared), at most three times.
    n_pass = max(1, args.warmup + args.steps)
    budget_s = 75.0
    n_thr = os.cpu_count() or 8
    cap = int(max(300, min(30000, 5500 * (budget_s / (3.0 * n_pass)) ** 0.5)))   # ~5e6 pairs/s per thread: 5500 rows = 3 s
    sample, desc = cpu_sample(inp, row_fraction=min(0.02, n_thr * cap / max(inp.n_rows, 1)), cap_rows=cap)
    for _ in range(3):
        t_cal = pyoracle.run(p, sample.c_ptr(), threads=0)["seconds_join"]
        if t_cal * n_pass <= 1.6 * budget_s or cap <= 300:
            break
        cap = int(max(300, cap * max(0.2, min(0.8, (budget_s / (t_cal * n_pass)) ** 0.5))))
        sample, desc = cpu_sample(inp, row_fraction=min(0.02, n_thr * cap / max(inp.n_rows, 1)), cap_rows=cap)
    # the bias of THIS rule, measured where the whole workload is affordable
    side_sample, side_desc = cpu_sample(side, row_fraction=min(0.10, 4 * n_thr * cap / max(side.n_rows, 1)), cap_rows=cap)
    ss = pyoracle.run(p, side_sample.c_ptr(), threads=0)
    full_v, samp_v = full["pairs_scored"] / full["seconds_join"], ss["pairs_scored"] / ss["seconds_join"]
    del side, side_sample
    times, pairs, threads = [], 0, 0
    for i in range(args.warmup + args.steps):
        r = pyoracle.run(p, sample.c_ptr(), threads=0)
        if i >= args.warmup:
            times.append(r["seconds_join"])
        pairs, threads = r["pairs_scored"], r["threads"]
    sec = sum(times) / max(len(times), 1)
    val = pairs / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 popc + f64 score",
        "data": "synthetic",
        "config": {"workload": WORKLOADS[MAIN_CONFIG], "timed": "CPU restatement of join_exacts (oracle port of the Rust reference; rustc/cargo absent), all host threads",
                   "sample": desc},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": desc,
                         "sample_bias": {"measured_on": WORKLOADS[SIDE_CONFIG], "whole_workload_pairs_per_s": full_v, "whole_workload_seconds": full["seconds_join"],
                                         "same_rule_sample_pairs_per_s": samp_v, "sample": side_desc, "full_over_sample": full_v / samp_v}},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------------
def pinned_arena(inp, torch):
    """Every input array moved once, outside the timed regions, into ONE page-locked allocation, the way the Rust shim flattens
    its records; the library then uploads it with a single DMA (ejb_set_input_arena)."""
    import enclone_ranger_b200 as E
    total = sum((arr.nbytes + 255) & ~255 for arr in inp.a.values())
    arena = torch.empty(total + 256, dtype=torch.uint8).pin_memory()
    arena_np = arena.numpy()
    base_off = (-arena_np.ctypes.data) & 255
    off, views = base_off, {}
    for name, arr in inp.a.items():
        view = arena_np[off:off + arr.nbytes].view(arr.dtype).reshape(arr.shape)
        view[...] = arr
        views[name] = view
        off += (arr.nbytes + 255) & ~255
    out = E.JoinInput(**views)
    assert all(out.a[n].ctypes.data == views[n].ctypes.data for n in views if views[n].size)
    return out, arena, arena_np.ctypes.data + base_off, off - base_off


def share_input(inp, rank, world, dist, tag):
    """rank 0 holds `inp`; every rank returns its own copy (through /dev/shm: one host process would hold the records once,
    here every rank process needs them in its own page-locked memory)."""
    import enclone_ranger_b200 as E
    from enclone_ranger_b200.inputs import _FIELDS
    d = f"/dev/shm/ejb_bench_{tag}"
    if rank == 0:
        os.makedirs(d, exist_ok=True)
        for name, arr in inp.a.items():
            np.save(os.path.join(d, name + ".npy"), arr)
    dist.barrier()
    if rank != 0:
        inp = E.JoinInput(**{name: np.load(os.path.join(d, name + ".npy"), mmap_mode="r") for name in _FIELDS})
    return inp, d


def measure_workload(E, torch, dist, args, cfg_index, cells, rank, world, local_rank, dev, peaks, with_cpu_baseline):
    """Device-resident and end-to-end arms of one workload; returns the result dict on rank 0 (None elsewhere)."""
    t_gen = time.time()
    inp_ascii = None
    if rank == 0:
        rep = E.generate(cfg_index, n_cells=cells)
        inp_ascii = rep.to_join_input()
        rep.close()
        # host buffers as the Rust shim hands them over: contigs and CDR3s 2-bit packed (info[k].tigsp) where they are pure
        # ACGT without indel marks, ASCII otherwise
        inp = inp_ascii if args.ascii else inp_ascii.to_packed().to_packed_cdr3()
        raw_tig_bytes = int(inp_ascii.a["tig_bytes"].size)
    else:
        inp, raw_tig_bytes = None, 0
    shm_dir = None
    if world > 1:
        inp, shm_dir = share_input(inp, rank, world, dist, os.environ.get("MASTER_PORT", "0"))
    t_gen = time.time() - t_gen
    params = E.default_params(True)
    inp, arena, arena_addr, arena_len = pinned_arena(inp, torch)
    if world > 1:
        dist.barrier()
        if rank == 0:
            import shutil
            shutil.rmtree(shm_dir, ignore_errors=True)
    ctx = E.JoinContext(device=local_rank, rank=rank, world=world)
    ctx.set_input_arena(arena_addr, arena_len)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    merged = {}

    def gather_and_merge():
        """per-rank partial forests -> rank 0's merge buffers over NCCL (send / recv straight into slices), merge on the device.
        Returns (gather ms on the device clock, merge ms) of this rank."""
        if world == 1:
            return 0.0, 0.0
        n, e, t = ctx.part_forest()
        sizes_t = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes_t, torch.tensor([n], dtype=torch.int64, device=dev))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sizes = sizes_t.tolist()
        total = int(sum(sizes))
        ops = []
        if rank == 0:
            ew, tw = ctx.merge_buffers(total)
            ew[:n].copy_(e)
            tw[:32 * n].copy_(t)
            off = n
            for r in range(1, world):
                if sizes[r]:
                    ops.append(dist.P2POp(dist.irecv, ew[off:off + sizes[r]], r))
                    ops.append(dist.P2POp(dist.irecv, tw[32 * off:32 * (off + sizes[r])], r))
                off += sizes[r]
        elif n:
            ops.append(dist.P2POp(dist.isend, e, 0))
            ops.append(dist.P2POp(dist.isend, t, 0))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        e1.record()
        torch.cuda.synchronize()
        merge_ms = 0.0
        if rank == 0:
            ms = ctx.merge_parts(total)
            merge_ms = ms["ms_merge"]
            merged.update({k: ms[k] for k in ("joins", "forest_edges", "two_cell_deleted")})
        return e0.elapsed_time(e1), merge_ms

    # ---------------- device-resident arm: value ----------------
    ctx.upload(params, inp)

    def step_resident():
        st = ctx.run()
        g_ms, m_ms = gather_and_merge()
        return st, st["ms_device_total"] + g_ms + m_ms, g_ms, m_ms

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    dev_ms, launches, stats, g_all, m_all = [], 0, None, [], []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        stats, ms, g_ms, m_ms = step_resident()
        dev_ms.append(ms)
        g_all.append(g_ms)
        m_all.append(m_ms)
        launches += stats["kernel_launches"]
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3 / args.steps
    clocks = sampler.stop() if sampler else None
    join_ms = [m - g - x for m, g, x in zip(dev_ms, g_all, m_all)]
    if world > 1:
        # the critical path of a step: the slowest rank's join, then rank 0's gather (it starts when the last partial forest is
        # ready) and merge.  Per step: max over ranks of the join + rank 0's gather + merge.
        tj = torch.tensor(join_ms, dtype=torch.float64, device=dev)
        dist.all_reduce(tj, op=dist.ReduceOp.MAX)
        tgm = torch.tensor([g + x for g, x in zip(g_all, m_all)], dtype=torch.float64, device=dev)
        dist.broadcast(tgm, 0)
        dev_ms = (tj + tgm).tolist()
    t_local = sum(dev_ms) / len(dev_ms)
    stats = dict(stats)
    stats.update(merged)
    stats["ms_merge"] = sum(m_all) / len(m_all)
    pairs = stats["pairs_scored"]
    per_rank = None
    if world > 1:
        t = torch.tensor([t_local, wall_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step, wall_ms = float(t[0].item()), float(t[1].item())
        ln = torch.tensor([launches, pairs], dtype=torch.int64, device=dev)
        dist.all_reduce(ln)
        launches, pairs = int(ln[0].item()), int(ln[1].item())
        keys = ("pairs_scored", "pairs_evaluated", "pairs_distance_computed", "stage1_candidates", "stage2_survivors", "pass_edges", "forest_edges", "rounds",
                "overflow_retries", "ms_pack", "ms_stage1", "ms_stage2", "ms_stage2a", "ms_forest", "ms_device_total")
        pr = torch.zeros(world * (2 + len(keys)), dtype=torch.float64, device=dev)
        mine = torch.tensor([sum(g_all) / len(g_all), sum(m_all) / len(m_all)] + [float(stats[k]) for k in keys], dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(pr, mine)
        pr = pr.reshape(world, -1).tolist()
        per_rank = [dict({"gather_ms": round(x[0], 3), "merge_ms": round(x[1], 3)}, **{k: round(v, 3) for k, v in zip(keys, x[2:])}) for x in pr]
    else:
        t_step = t_local

    # ---------------- end-to-end arm: host buffers through the public API ----------------
    def step_e2e():
        if world == 1:
            return ctx.join(params, inp, copy=False)   # zero-copy views into the library's pinned result arena
        # the input crosses PCIe once: every rank uploads its 1/N slice of its pinned arena, the slices are all-gathered over NVLink
        ctx.upload_slice(params, inp, rank, world)
        arena_dev, sl = ctx.device_arena()
        dist.all_gather_into_tensor(arena_dev, arena_dev[rank * sl:(rank + 1) * sl].clone())
        torch.cuda.synchronize()
        ctx.run()
        gather_and_merge()
        return ctx.download(copy=False) if rank == 0 else None

    step_e2e()
    barrier()
    e2e_times = []
    res = None
    for _ in range(args.steps):
        t0 = time.perf_counter()
        res = step_e2e()
        if world > 1:
            dist.barrier()
        e2e_times.append((time.perf_counter() - t0) * 1e3)
    barrier()
    e2e_ms = sum(e2e_times) / len(e2e_times)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    out = None
    if rank == 0:
        e2e_stats = res.stats
        # the committed full-size oracle digest of this workload (tests/golden/digests.json): the benchmarked result IS the
        # reference's result
        digest_ok = None
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
            from make_digests import ARRAYS, digest_of
            want = json.load(open(os.path.join(ROOT, "tests", "golden", "digests.json"))).get(f"config{cfg_index}" if cells is None else None)
            if want is not None:
                got = digest_of(res, inp.n_rows)
                digest_ok = all(got[f] == want[f] for f in ARRAYS) and got["n_edges"] == want["n_edges"]
        except Exception as ex:   # noqa: BLE001
            digest_ok = f"not checked: {ex}"
        p1_avg, p2_avg, dom_pairs, packed_bytes = algorithmic_popc(inp_ascii)
        s1_ms, s2a_ms = stats["ms_stage1"], stats["ms_stage2a"]
        peak = peaks["popc_per_s"]
        alg_popc = pairs * p1_avg + sum_stat(per_rank, stats, "stage1_candidates") * p2_avg
        kernels = {
            "k_stage1": {"ms_per_step": s1_ms, "launches_per_step": stats["stage1_launches"], "pairs_of_staged_tiles": stats["pairs_evaluated"],
                         "pairs_distance_computed": stats["pairs_distance_computed"], "pairs_distance_full": stats["pairs_distance_full"],
                         "popc_algorithmic_per_pair": p1_avg, "popc_executed": stats["popc_executed_stage1"],
                         "frac_effective": stats["pairs_evaluated"] * p1_avg / max(s1_ms, 1e-9) / 1e-3 / peak,
                         "frac_executed": stats["popc_executed_stage1"] / max(s1_ms, 1e-9) / 1e-3 / peak,
                         "note": "effective = algorithmic popc of every pair of the staged tiles (skipped 16x16 groups included) / time / peak; "
                                 "executed = popc instructions stage 1 really issued (lane granular) / time / peak — compare with ncu smsp__inst_executed_pipe_xu"},
            "k_stage2a": {"ms_per_step": s2a_ms, "launches_per_step": stats["stage2a_launches"], "candidates": stats["stage1_candidates"], "popc_per_candidate": p2_avg,
                          "frac_effective": stats["stage1_candidates"] * p2_avg / max(s2a_ms, 1e-9) / 1e-3 / peak},
        }
        if world > 1:
            kernels["note"] = "per-kernel figures are rank 0's share"
        dom = "k_stage2a" if s2a_ms >= s1_ms else "k_stage1"
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            hbm_src = "MEASURED_PEAKS.json"
        except Exception:   # noqa: BLE001
            hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
        traffic, traffic_src = None, None
        try:
            nt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")))
            traffic, traffic_src = nt[dom]["dram_bytes_per_launch"], "profiles/ncu_traffic_r2.json (ncu --set full of configs[1]; dram__bytes_read.sum + dram__bytes_write.sum per launch)"
        except Exception:   # noqa: BLE001
            pass
        alg_bytes = packed_bytes + 8.0 * stats["joins"] + 4.0 * inp.n_rows
        whole_frac = alg_popc / (t_step * 1e-3) / (peak * world)
        out = {
            "value": pairs / (t_step / 1e3), "ms_per_step": t_step, "ms_per_step_spread": spread(dev_ms), "wall_ms_per_step": wall_ms,
            "workload": WORKLOADS[cfg_index] if cells is None else f"DEBUG {cells}-cell variant of configs[{cfg_index}]",
            "cells": int(inp_ascii.n_cells), "rows": int(inp.n_rows), "pairs_scored": int(pairs), "pairs_lens_buckets": int(stats["pairs_lens"]),
            "sub_buckets": int(stats["n_subbuckets"]), "generate_s": t_gen, "per_rank": per_rank, "clocks": clocks, "gpu_launches": int(launches),
            "contig_encoding": "ASCII" if args.ascii else f"2-bit packed contigs and CDR3s ({raw_tig_bytes / 1e6:.0f} MB ASCII contigs -> {inp.a['tig2_words'].nbytes / 1e6:.0f} MB packed + {inp.a['tig_bytes'].nbytes / 1e6:.0f} MB ASCII)",
            "oracle_digest_match": digest_ok,
            "e2e": {"value": pairs / (e2e_ms / 1e3), "unit": "pairs/s", "ms_per_step": e2e_ms, "ms_per_step_spread": spread(e2e_times),
                    "h2d_bytes_per_step": int(e2e_stats["h2d_bytes"]) * world, "d2h_bytes_per_step": int(e2e_stats["d2h_bytes"]),
                    "pinned_host_buffers": "one arena per rank (ejb_set_input_arena)",
                    "breakdown_ms": {k: e2e_stats[k] for k in ("ms_h2d", "ms_device_total", "ms_d2h", "ms_replay")},
                    "note": "H2D of 1/N of the input per rank + NVLink all-gather of the slices + partition + run + NCCL gather + merge + D2H + EquivRel replay" if world > 1 else "H2D + run + D2H + EquivRel replay"},
            "roofline": {"bound": "popc", "kernel": dom, "achieved": kernels[dom]["frac_effective"] * peak / 1e9, "peak": peak / 1e9, "unit": "Gpopc/s",
                         "frac": kernels[dom]["frac_effective"], "frac_executed": kernels[dom].get("frac_executed"),
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "measured in this job (ejb_measure_peaks: dependency-free __popc loop, 16 lanes/clk/SM)",
                         "lop3_peak_gops": peaks["lop3_per_s"] / 1e9, "kernels": kernels,
                         "whole_join": {"popc_per_pair": alg_popc / max(pairs, 1), "algorithmic_popc": alg_popc, "n_gpus": world,
                                        "frac": whole_frac, "note": "algorithmic popc of the whole join / step time / (peak x n_gpus)"},
                         "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (t_step / 1e3) / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src}},
            "stage_ms": {k: stats[k] for k in ("ms_pack", "ms_stage1", "ms_stage2", "ms_stage2a", "ms_forest", "ms_finalize", "ms_merge")},
            "counts": {k: int(stats[k]) for k in ("pairs_evaluated", "pairs_distance_computed", "pairs_distance_full", "popc_executed_stage1", "stage1_candidates",
                                                  "stage1_flag_skips", "stage1_ref_kills", "stage1_donor_kills", "stage2_two_ref", "stage2_memo_dead", "stage2_degr_dead",
                                                  "stage2_slow", "stage2_survivors", "pass_edges", "forest_edges", "joins", "rounds", "host_syncs",
                                                  "boruvka_iterations", "kernel_launches")},
            "create_ms": ctx.create_ms(),
        }
        if with_cpu_baseline:
            from oracle import pyoracle   # CPU baseline leg only: the checker timed as a baseline, never on the product path
            empty = E.JoinInput()
            pyoracle.run(params, empty.c_ptr(), threads=0)
            r = pyoracle.run(params, inp_ascii.c_ptr(), threads=0)
            out["cpu_baseline"] = {"value": r["pairs_scored"] / r["seconds_join"], "unit": "pairs/s", "cores": int(r["threads"]), "kind": "port",
                                   "sample": f"the WHOLE workload ({WORKLOADS[cfg_index]}), one pass", "seconds": r["seconds_join"],
                                   "join_one_calls": int(r["join_one_calls"]), "pairs": int(r["pairs_scored"]),
                                   "edges_equal_gpu": bool(len(r["edge_k1"]) == len(res.edge_k1) and np.array_equal(r["edge_k1"], res.edge_k1) and np.array_equal(r["edge_k2"], res.edge_k2))}
    ctx.close()
    del arena
    return out


def sum_stat(per_rank, stats, key):
    return sum(r[key] for r in per_rank) if per_rank else stats[key]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=None, help="override the repertoire size of the main workload (debug only; the number is then not the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the configs[1] side measurement of the N = 1 line")
    ap.add_argument("--ascii", action="store_true", help="hand every contig / CDR3 over as ASCII (info[k].tigs) instead of 2-bit packed (tigsp)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import enclone_ranger_b200 as E
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the join has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = None
    if rank == 0:
        pctx = E.JoinContext(device=local_rank)
        peaks = pctx.measure_peaks()
        pctx.close()
    main_res = measure_workload(E, torch, dist, args, MAIN_CONFIG, args.cells, rank, world, local_rank, dev, peaks, with_cpu_baseline=False)
    side_res = None
    if world == 1 and not args.no_side:
        side_res = measure_workload(E, torch, dist, args, SIDE_CONFIG, None, rank, world, local_rank, dev, peaks, with_cpu_baseline=not args.no_cpu_baseline)
    if rank == 0:
        m = main_res
        line = {
            "metric": METRIC, "value": m["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["ms_per_step"], "ms_per_step_spread": m["ms_per_step_spread"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 popc + f64 score", "data": "synthetic",
            "config": {"workload": m["workload"], "cells": m["cells"], "rows": m["rows"], "pairs_scored": m["pairs_scored"], "pairs_lens_buckets": m["pairs_lens_buckets"],
                       "sub_buckets": m["sub_buckets"],
                       "parallelism": (f"strong scaling: the same repertoire sharded over {world} ranks by the library (sub-buckets by LPT, the heaviest cut into row ranges; "
                                       "partition inside the timed region), partial forests gathered over NCCL and merged on rank 0's GPU") if world > 1 else "single GPU",
                       "l2": "inputs (GBs of raw input + packed row store) exceed the 126 MB L2; no flush needed",
                       "timing": ("CUDA events on the library's launching stream (ejb_stats.ms_device_total)" if world == 1 else
                                  "device clock: CUDA events of the join on the library's stream + CUDA events around the NCCL gather on torch's stream + CUDA events of the merge; per step = max over ranks of the join + rank 0's gather + merge (the critical path)") +
                                 "; wall clock of the same K steps (barrier + synchronize on both sides) in wall_ms_per_step",
                       "wall_ms_per_step": m["wall_ms_per_step"], "generate_s": m["generate_s"], "per_rank": m["per_rank"], "contig_encoding": m["contig_encoding"],
                       "oracle_digest_match": m["oracle_digest_match"], "create_ms": m["create_ms"]},
            "clocks": m["clocks"], "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "roofline": m["roofline"], "stage_ms": m["stage_ms"], "counts": m["counts"],
        }
        if side_res is not None:
            line["configs1"] = side_res
            if "cpu_baseline" in side_res:
                line["cpu_baseline"] = side_res["cpu_baseline"]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
