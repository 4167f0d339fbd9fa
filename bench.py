#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 ZPAQ block codec.

Metric (BASELINE.json): compress & decompress input MB/s per -mN at 1/2/4/8 B200 vs host CPU.
Workload at N=1 (BASELINE.json configs[1]): -m2 (ICM + 2 x ISSE), 1 GiB of synthetic text in
1024 independent 1 MiB blocks.  With N GPUs every rank gets its own 1 GiB (weak scaling): blocks
are independent, so ranks share nothing on the data path; torch.distributed is only used for the
barrier and the max-over-ranks of the timed region.

One step = compress every block, then decompress every block (one pass of the hot path both
ways).  value = input bytes / (t_compress + t_decompress), inputs resident in HBM.  e2e = the same
through the host-buffer C-ABI calls (zpaqgpu_compress_blocks / zpaqgpu_decompress_archive) with the
host<->device copies inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--level L] [--blocks B] [--block-kib S]
  python bench.py --impl reference ...   # the reference algorithm on the host cores (CPU oracle)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "compress & decompress input MB/s per -mN at 1/2/4/8 B200 vs host CPU"
# integer ops per input byte for m1..m5 (counted from the reference source)
W_OPS = {1: 1010, 2: 1430, 3: 2300, 4: 3000, 5: 3870}
N_HT = {1: 2, 2: 3, 3: 5, 4: 6, 5: 8}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--blocks", type=int, default=1024)
    ap.add_argument("--block-kib", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured"
    except Exception:
        return 6650.0, 1965.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx = float(parts[2])
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = mx
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def oracle_run(level, data, n_blocks, block_bytes, threads):
    """Compress then decompress n_blocks with the CPU oracle on `threads` host threads.
    Returns (t_compress, t_decompress, archive_bytes)."""
    import numpy as np
    import oracle_binding as ob
    L = ob.lib()
    off = (C.c_uint64 * (n_blocks + 1))(*[i * block_bytes for i in range(n_blocks + 1)])
    cap = n_blocks * (block_bytes + block_bytes // 4 + 4096)
    out = np.empty(cap, dtype=np.uint8)
    out_off = (C.c_uint64 * (n_blocks + 1))()
    need = C.c_uint64(0)
    src = np.ascontiguousarray(data[:n_blocks * block_bytes])
    t0 = time.perf_counter()
    rc = L.zo_compress_blocks_mt(level, src.ctypes.data, off, n_blocks, out.ctypes.data, cap, out_off,
                                 C.byref(need), threads)
    t1 = time.perf_counter()
    assert rc == 0
    back = np.empty(n_blocks * block_bytes + 16, dtype=np.uint8)
    back_off = (C.c_uint64 * (n_blocks + 1))()
    rc = L.zo_decompress_blocks_mt(out.ctypes.data, out_off, n_blocks, back.ctypes.data, len(back), back_off,
                                   C.byref(need), threads)
    t2 = time.perf_counter()
    assert rc == 0 and need.value == n_blocks * block_bytes
    assert bytes(back[:need.value]) == bytes(src), "oracle round trip failed"
    return t1 - t0, t2 - t1, int(out_off[n_blocks]), out, out_off


def cpu_sample_blocks(level, cores, block_bytes):
    # about 10-30 s of CPU work on all cores: the oracle codes ~0.6-1.7 MB/s per thread each way
    per_thread = {1: 6, 2: 6, 3: 4, 4: 3, 5: 2}.get(level, 4)
    scale = max(1, (1 << 20) // block_bytes)
    return max(cores * per_thread * scale, 1)


def run_reference(args):
    """--impl reference: the reference's algorithm (CPU oracle: the V source cannot be compiled in
    this image, there is no V toolchain) on all host threads, a bounded sample per step."""
    import datagen
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    block_bytes = args.block_kib * 1024
    nb = min(args.blocks, max(cores, cpu_sample_blocks(args.level, cores, block_bytes) // 3))
    data = datagen.text_stream(nb * block_bytes)
    for _ in range(args.warmup):
        oracle_run(args.level, data, min(nb, cores), block_bytes, cores)
    tc = td = 0.0
    arc = 0
    for _ in range(args.steps):
        a, b, arc, _, _ = oracle_run(args.level, data, nb, block_bytes, cores)
        tc += a
        td += b
    total = nb * block_bytes * args.steps
    value = total / (tc + td) / 1e6
    sample = "%d x %d KiB text blocks per step (of %d), compress+decompress, %d threads" % (
        nb, args.block_kib, args.blocks, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round((tc + td) / args.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(args),
        "compress_mb_s": round(total / tc / 1e6, 3), "decompress_mb_s": round(total / td / 1e6, 3),
        "ratio": round(arc / (nb * block_bytes), 4),
        "cpu_baseline": {"value": round(value, 3), "unit": "MB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def config_dict(args):
    return {"workload": "-m%d, %d x %d KiB independent blocks of synthetic text per GPU (BASELINE.json configs[1])"
                        % (args.level, args.blocks, args.block_kib),
            "level": args.level, "blocks_per_gpu": args.blocks, "block_bytes": args.block_kib * 1024,
            "l2": "inputs (%d MiB per GPU) exceed the 126 MB L2, no flush needed" % (args.blocks * args.block_kib // 1024),
            "parallelism": "one ZPAQ block per warp; disjoint block ranges per GPU, no collective"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import datagen
    import zpaq_v_b200 as z
    from zpaq_v_b200 import binding as zb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libzpaqgpu has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner (printed by the library when the
        # first communicator comes up) is kept off fd 1
        os.environ["NCCL_DEBUG"] = "WARN"
        sys.stdout.flush()
        saved, null = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(null, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(null)

    level, nb, bb = args.level, args.blocks, args.block_kib * 1024
    total = nb * bb
    dev = torch.device("cuda", local)
    # every rank codes its own stretch of the text stream
    host_np = datagen.text_stream(total, first=rank * total)
    host_in = torch.from_numpy(host_np.copy()).pin_memory()
    d_in = host_in.to(dev, non_blocking=False)
    cap = total + total // 4 + 4096 * nb
    d_arc = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_arc_off = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
    d_plain = torch.empty(total, dtype=torch.uint8, device=dev)
    d_plain_len = torch.zeros(nb, dtype=torch.int64, device=dev)
    in_off = (C.c_uint64 * (nb + 1))(*[i * bb for i in range(nb + 1)])

    ctx = z.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    L = zb.lib()
    launches = {"n": 0}
    kern_ms = {"enc": 0.0, "dec": 0.0, "enc_n": 0, "dec_n": 0}

    def step_device(timed):
        """compress + decompress with inputs resident in HBM; returns (t_c, t_d) in seconds"""
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tot = C.c_uint64(0)
        e[0].record()
        ctx._check(L.zpaqgpu_compress_blocks_dev(ctx._h, level, d_in.data_ptr(), None, in_off, nb, d_arc.data_ptr(),
                                                 cap, d_arc_off.data_ptr(), C.byref(tot)))
        e[1].record()
        st_c = ctx.stats()
        arc_off_h = d_arc_off.cpu().numpy().astype(np.uint64)
        arc_off_c = (C.c_uint64 * (nb + 1))(*arc_off_h.tolist())
        bad = C.c_int(0)
        e1b = torch.cuda.Event(enable_timing=True)
        e1b.record()
        ctx._check(L.zpaqgpu_decompress_blocks_dev(ctx._h, d_arc.data_ptr(), arc_off_c, nb, d_plain.data_ptr(), in_off,
                                                   d_plain_len.data_ptr(), C.byref(bad)))
        e[2].record()
        torch.cuda.synchronize()
        st_d = ctx.stats()
        if bad.value:
            raise SystemExit("decompression reported %d bad blocks" % bad.value)
        if timed:
            launches["n"] += st_c["launches"] + st_d["launches"]
            kern_ms["enc"] += st_c["codec_ms"]; kern_ms["enc_n"] += st_c["codec_launches"]
            kern_ms["dec"] += st_d["codec_ms"]; kern_ms["dec_n"] += st_d["codec_launches"]
        return e[0].elapsed_time(e[1]) / 1e3, e1b.elapsed_time(e[2]) / 1e3, int(tot.value), st_c, st_d

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device(False)
    if not torch.equal(d_plain, d_in):
        raise SystemExit("round trip mismatch on the device path")
    clocks = ClockSampler(local if os.environ.get("CUDA_VISIBLE_DEVICES") is None else
                          int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local]))
    barrier()
    if rank == 0:
        clocks.start()
    tc = td = 0.0
    arc_total = 0
    st_c = st_d = None
    for _ in range(args.steps):
        a, b, arc_total, st_c, st_d = step_device(True)
        tc += a
        td += b
    barrier()
    clk = clocks.stop() if rank == 0 else {}
    t_all = torch.tensor([tc + td, tc, td], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    t_sum, t_c, t_d = [float(x) for x in t_all.cpu()]
    bytes_all = total * world * args.steps
    value = bytes_all / t_sum / 1e6

    # parity sample against the oracle (outside the timed region): first blocks byte-identical
    parity = None
    if rank == 0:
        import oracle_binding as ob
        arc_off_h = d_arc_off.cpu().numpy()
        ok = True
        for b in (0, nb - 1):
            want = ob.compress_block(level, host_np[b * bb:(b + 1) * bb].tobytes(), "", "%d bytes" % bb)
            got = bytes(d_arc[int(arc_off_h[b]):int(arc_off_h[b + 1])].cpu().numpy())
            ok = ok and got == want
        parity = bool(ok)

    # ---- end to end through the host-buffer C ABI ----
    e2e = None
    if not args.no_e2e:
        host_arc = torch.empty(cap, dtype=torch.uint8).pin_memory()
        host_out = torch.empty(total, dtype=torch.uint8).pin_memory()
        out_off = (C.c_uint64 * (nb + 1))()
        need = C.c_uint64(0)
        comments = (C.c_char_p * nb)(*[b"%d bytes" % bb for _ in range(nb)])
        nseg = C.c_int(0)
        segs = (zb.Segment * (nb + 8))()

        def step_host():
            t0 = time.perf_counter()
            ctx._check(L.zpaqgpu_compress_blocks(ctx._h, level, host_in.data_ptr(), in_off, nb, None, comments,
                                                 host_arc.data_ptr(), cap, out_off, C.byref(need)))
            t1 = time.perf_counter()
            arc_len = int(out_off[nb])
            rc = L.zpaqgpu_decompress_archive(ctx._h, host_arc.data_ptr(), arc_len, host_out.data_ptr(), total,
                                              C.byref(need), segs, nb + 8, C.byref(nseg))
            ctx._check(rc)
            t2 = time.perf_counter()
            return t1 - t0, t2 - t1, arc_len

        e_steps = max(1, min(args.steps, 2))
        step_host()
        if not torch.equal(host_out, host_in):
            raise SystemExit("round trip mismatch on the host path")
        barrier()
        hc = hd = 0.0
        arc_len = 0
        for _ in range(e_steps):
            a, b, arc_len = step_host()
            hc += a
            hd += b
        barrier()
        t_h = torch.tensor([hc + hd], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_h, op=dist.ReduceOp.MAX)
        e2e = {"value": round(total * world * e_steps / float(t_h.item()) / 1e6, 3), "unit": "MB/s",
               "h2d_bytes_per_step": int(total + arc_len), "d2h_bytes_per_step": int(arc_len + total),
               "compress_mb_s": round(total * e_steps / hc / 1e6, 3),
               "decompress_mb_s": round(total * e_steps / hd / 1e6, 3), "steps": e_steps}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----
    # The decode kernel takes the larger share of the step; the encoder is reported beside it.
    hbm_peak, sm_max, peak_kind = peaks()
    ratio = arc_total / total
    n_ht = N_HT.get(level, 3)
    # algorithmic bytes per input byte the codec kernel asks of the memory system:
    # plaintext (1) + coded bytes (ratio) + per nibble and hash-table component one 64-byte probe line
    # read and one 16-byte slot write-back (2 nibbles per byte).  The tree decoder (-m1..-m3) requests
    # the four possible lines of the next nibble two bits early: four line reads per probe instead of one.
    tree_dec = level <= 3 and os.environ.get("ZPAQGPU_DECODER", "tree") != "serial"
    spec = 4 if (tree_dec and os.environ.get("ZPAQGPU_SPEC_PROBE", "1") != "0") else 1
    alg_enc = 1.0 + ratio + 2 * n_ht * (64 + 16)
    alg_dec = 1.0 + ratio + 2 * n_ht * (64 * spec + 16)
    enc_ms = kern_ms["enc"] / max(1, kern_ms["enc_n"])
    dec_ms = kern_ms["dec"] / max(1, kern_ms["dec_n"])
    launch_bytes = total / max(1, st_d["waves"])
    f_clk = (clk.get("sm_mhz") or sm_max) * 1e6
    issue_peak = 148 * 128 * f_clk
    traffic = {"encode": None, "decode": None}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if level == 2:
            traffic = {k: round(tj[k]["dram_bytes_per_input_byte"] * launch_bytes) for k in ("encode", "decode")}
    except Exception:
        pass

    def roof(ms, name, key, alg):
        ach = launch_bytes * alg / (ms / 1e3) / 1e9
        return {"kernel": name, "achieved": round(ach, 2), "frac": round(ach / hbm_peak, 5), "kernel_ms": round(ms, 3),
                "traffic": traffic[key], "algorithmic_bytes_per_input_byte": round(alg, 2),
                "issue_frac": round(launch_bytes / (ms / 1e3) * W_OPS.get(level, 0) / issue_peak, 5)}

    n_isse = N_HT.get(level, 3) - 1
    mix = "true" if level >= 4 else "false"
    dec_r = roof(dec_ms, "k_decode_chain<%d,%s,%s>" % (n_isse, mix, "true" if tree_dec else "false"), "decode", alg_dec)
    enc_r = roof(enc_ms, "k_encode_pipe3<%d,%s>" % (n_isse, mix), "encode", alg_enc)
    roofline = {
        "bound": "hbm", "kernel": dec_r["kernel"], "achieved": dec_r["achieved"], "peak": hbm_peak, "unit": "GB/s",
        "frac": dec_r["frac"], "peak_source": peak_kind, "traffic": dec_r["traffic"],
        "algorithmic_bytes_per_input_byte": dec_r["algorithmic_bytes_per_input_byte"], "units_per_launch": int(launch_bytes),
        "kernel_ms": dec_r["kernel_ms"], "encode": enc_r,
        "note": "bit-serial integer chain: the binding limit is the dependent-instruction latency of one warp per "
                "block, not HBM; issue_roofline = W(L) ops/byte x bytes/s / (148 SM x 128 lanes x f_clk)",
        "issue_roofline": {"ops_per_input_byte": W_OPS.get(level), "decode_frac": dec_r["issue_frac"],
                           "encode_frac": enc_r["issue_frac"], "peak_ops_per_s": issue_peak, "sm_mhz": f_clk / 1e6},
    }

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_s = min(nb, cpu_sample_blocks(level, cores, bb))
        a, b, _, _, _ = oracle_run(level, host_np, n_s, bb, cores)
        cpu = {"value": round(n_s * bb / (a + b) / 1e6, 3), "unit": "MB/s", "cores": cores, "kind": "port",
               "sample": "first %d of %d blocks (%d KiB each), compress+decompress, CPU oracle (C restatement of the "
                         "V reference) on %d threads" % (n_s, nb, args.block_kib, cores),
               "compress_mb_s": round(n_s * bb / a / 1e6, 3), "decompress_mb_s": round(n_s * bb / b / 1e6, 3)}

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(t_sum / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(args),
        "compress_mb_s": round(bytes_all / t_c / 1e6, 3), "decompress_mb_s": round(bytes_all / t_d / 1e6, 3),
        "ratio": round(ratio, 4), "byte_identical_to_oracle": parity,
        "e2e": e2e, "gpu_launches": launches["n"], "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
        "stats": {"compress": st_c, "decompress": st_d},
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
