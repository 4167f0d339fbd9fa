#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 ZPAQ block codec.

Metric (BASELINE.json): compress & decompress input MB/s per -mN at 1/2/4/8 B200 vs host CPU.
Workload of `value` (BASELINE.json configs[1]): -m2 (ICM + 2 x ISSE), 1 GiB of synthetic text in
1024 independent 1 MiB blocks.  With N GPUs every rank gets its own 1 GiB (weak scaling): blocks
are independent, so ranks share nothing on the data path; torch.distributed is only used for the
barrier and the max-over-ranks of the timed region.

One step = compress every block, then decompress every block (one pass of the hot path both
ways).  value = input bytes / (t_compress + t_decompress), inputs resident in HBM.  e2e = the same
through the host-buffer C-ABI calls (zpaqgpu_compress_blocks / zpaqgpu_decompress_archive) with the
host<->device copies inside the timed region.

After the headline the same line gets a `per_level` block: one pass each of the other BASELINE.json
configurations on the same GPU(s) -- -m1 and -m4 at 1024 x 1 MiB text, cfg 3 (-m3, 8 GiB mixed text /
random / structured in 1 MiB blocks, the 8192 blocks split over the ranks: strong scaling), cfg 4 (-m5,
1024 x 4 MiB text, decompression is the quoted figure), cfg 5 (jidac add of a 10 000-file tree, rank 0)
-- each with kernel times from CUDA events, the issue-roofline fraction, table mode, waves and a
byte-for-byte comparison of at least 8 blocks with the CPU oracle outside the timed region.

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
N_SM, LANES = 148, 128


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
    ap.add_argument("--e2e-steps", type=int, default=5, help="cap of the end-to-end timed steps (stated in the line)")
    ap.add_argument("--per-level", default="1,3,4,5,jidac,generic", help="extra configurations after the headline")
    ap.add_argument("--no-per-level", action="store_true")
    ap.add_argument("--cfg3-blocks", type=int, default=8192, help="cfg 3: blocks of 1 MiB over ALL ranks")
    ap.add_argument("--cfg4-blocks", type=int, default=1024, help="cfg 4: blocks of 4 MiB per rank")
    ap.add_argument("--jidac-files", type=int, default=10000)
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


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


def oracle_blocks(level, src, offs, threads):
    """The CPU oracle over blocks src[offs[k]:offs[k+1]] on `threads` host threads: compress, then
    decompress.  Returns (t_compress, t_decompress, archive array, archive offsets)."""
    import numpy as np
    import oracle_binding as ob
    L = ob.lib()
    n = len(offs) - 1
    total = int(offs[-1] - offs[0])
    off = (C.c_uint64 * (n + 1))(*[int(x) for x in offs])
    cap = total + total // 2 + 4096 * n
    out = np.empty(cap, dtype=np.uint8)
    out_off = (C.c_uint64 * (n + 1))()
    need = C.c_uint64(0)
    src = np.ascontiguousarray(src)
    t0 = time.perf_counter()
    rc = L.zo_compress_blocks_mt(level, src.ctypes.data, off, n, out.ctypes.data, cap, out_off, C.byref(need), threads)
    t1 = time.perf_counter()
    assert rc == 0, "oracle compress failed (%d)" % rc
    back = np.empty(total + 16, dtype=np.uint8)
    back_off = (C.c_uint64 * (n + 1))()
    rc = L.zo_decompress_blocks_mt(out.ctypes.data, out_off, n, back.ctypes.data, len(back), back_off,
                                   C.byref(need), threads)
    t2 = time.perf_counter()
    assert rc == 0 and need.value == total
    assert bytes(back[:total]) == bytes(src[int(offs[0]):int(offs[-1])]), "oracle round trip failed"
    return t1 - t0, t2 - t1, out, [int(out_off[k]) for k in range(n + 1)]


def oracle_run(level, data, n_blocks, block_bytes, threads):
    a, b, out, out_off = oracle_blocks(level, data[:n_blocks * block_bytes],
                                       [i * block_bytes for i in range(n_blocks + 1)], threads)
    return a, b, out_off[-1], out, out_off


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
            "l2": ("inputs (%d MiB per GPU) exceed the 126 MB L2, no flush needed" if args.blocks * args.block_kib * 1024 > 126e6
                   else "inputs (%d MiB per GPU) FIT the 126 MB L2 and nothing flushes it between steps: a smoke "
                        "configuration, not a bench line") % (args.blocks * args.block_kib // 1024),
            "parallelism": "one ZPAQ block per warp group; disjoint block ranges per GPU, no collective"}


def load_traffic():
    """DRAM bytes per launch of the codec kernels from the committed ncu pass on the bench configuration
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from profiles/r02_traffic_launches.csv)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class DeviceCodec:
    """compress_blocks_dev / decompress_blocks_dev over equal-sized blocks resident in HBM."""

    def __init__(self, z, zb, torch, dev, stream_ptr):
        self.z, self.zb, self.torch, self.dev = z, zb, torch, dev
        self.L = zb.lib()
        self.ctx = z.Context(dev.index)
        self.ctx.set_stream(stream_ptr)

    def close(self):
        self.ctx.close()

    def alloc(self, nb, bb, slack=4):
        t = self.torch
        total = nb * bb
        self.nb, self.bb, self.total = nb, bb, total
        self.cap = total + total // slack + 4096 * nb
        self.d_arc = t.empty(self.cap, dtype=t.uint8, device=self.dev)
        self.d_arc_off = t.zeros(nb + 1, dtype=t.int64, device=self.dev)
        self.d_plain = t.empty(total, dtype=t.uint8, device=self.dev)
        self.d_plain_len = t.zeros(nb, dtype=t.int64, device=self.dev)
        self.in_off = (C.c_uint64 * (nb + 1))(*[i * bb for i in range(nb + 1)])

    def presize(self, level, d_in):
        """The library's table buffers grow to what nb blocks need BEFORE the timed pass: one pass over nb blocks
        of 64 bytes each (tables are per block, whatever its size).  Without it the timed call pays for a
        cudaFree + cudaMalloc of tens of GB (-m1: 38.7 GB of dense tables; 0.3..1.5 s of a 1.6 s call)."""
        unit = 64
        if self.bb < unit:
            return
        self.step(level, d_in, in_off=(C.c_uint64 * (self.nb + 1))(*[i * unit for i in range(self.nb + 1)]))

    def step(self, level, d_in, in_off=None):
        """Returns (t_compress_s, t_decompress_s, archive_bytes, stats_c, stats_d); times from CUDA events
        on the stream the kernels run on."""
        import numpy as np
        t, ctx, L, nb = self.torch, self.ctx, self.L, self.nb
        in_off = self.in_off if in_off is None else in_off
        e = [t.cuda.Event(enable_timing=True) for _ in range(4)]
        tot = C.c_uint64(0)
        e[0].record()
        ctx._check(L.zpaqgpu_compress_blocks_dev(ctx._h, level, d_in.data_ptr(), None, in_off, nb,
                                                 self.d_arc.data_ptr(), self.cap, self.d_arc_off.data_ptr(),
                                                 C.byref(tot)))
        e[1].record()
        st_c = ctx.stats()
        arc_off_h = self.d_arc_off.cpu().numpy().astype(np.uint64)
        arc_off_c = (C.c_uint64 * (nb + 1))(*arc_off_h.tolist())
        bad = C.c_int(0)
        e[2].record()
        ctx._check(L.zpaqgpu_decompress_blocks_dev(ctx._h, self.d_arc.data_ptr(), arc_off_c, nb, self.d_plain.data_ptr(),
                                                   in_off, self.d_plain_len.data_ptr(), C.byref(bad)))
        e[3].record()
        t.cuda.synchronize()
        st_d = ctx.stats()
        if bad.value:
            raise SystemExit("decompression reported %d bad blocks" % bad.value)
        return e[0].elapsed_time(e[1]) / 1e3, e[2].elapsed_time(e[3]) / 1e3, int(tot.value), st_c, st_d

    def block_bytes_of(self, b):
        off = self.d_arc_off[b:b + 2].cpu().numpy()
        return bytes(self.d_arc[int(off[0]):int(off[1])].cpu().numpy())


def parity_blocks(codec, level, host_np, idx, cores):
    """Blocks idx of the last compress step, byte for byte against the CPU oracle (comment "<n> bytes",
    empty name: what compress_blocks_dev writes)."""
    import numpy as np
    import oracle_binding as ob
    bb = codec.bb
    if len(idx) <= 2:
        return all(codec.block_bytes_of(b) == ob.compress_block(level, host_np[b * bb:(b + 1) * bb].tobytes(), "",
                                                                  "%d bytes" % bb) for b in idx)
    sample = np.concatenate([host_np[b * bb:(b + 1) * bb] for b in idx])
    _, _, out, out_off = oracle_blocks(level, sample, [k * bb for k in range(len(idx) + 1)], min(cores, len(idx)))
    # zo_compress_blocks_mt writes the same framing (empty name, "<n> bytes")
    return all(codec.block_bytes_of(b) == bytes(out[out_off[k]:out_off[k + 1]]) for k, b in enumerate(idx))


def issue_frac(level, input_bytes, seconds, f_clk_hz):
    return input_bytes / seconds * W_OPS[level] / (N_SM * LANES * f_clk_hz)


def hbm_block(level, ratio, ws_bytes_per_block, paged, bb, input_bytes, seconds, hbm_peak):
    """algorithmic HBM bytes per input byte = 1 (read) + ratio (write) + table bytes cleared per
    block / block size; the upper bound adds every probe as a miss: 2 probes x n_ht x 64 B x 2 (read +
    write-back).  No speculative reads are counted."""
    alg = 1.0 + ratio + ws_bytes_per_block / bb
    upper = alg + 2 * N_HT[level] * 128
    ach = input_bytes * alg / seconds / 1e9
    return {"algorithmic_bytes_per_input_byte": round(alg, 3), "all_probes_miss_bytes_per_input_byte": round(upper, 3),
            "table_bytes_cleared_per_block": int(ws_bytes_per_block), "tables": "paged" if paged else "dense",
            "achieved": round(ach, 3), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 6),
            "frac_if_all_probes_miss": round(input_bytes * upper / seconds / 1e9 / hbm_peak, 6)}


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
        # rank 0 prints ONE JSON line on stdout: whatever NCCL prints while the communicator comes up is
        # kept off fd 1 (NCCL_DEBUG itself is left as the caller set it)
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
    cores = os.cpu_count() or 1
    hbm_peak, sm_max, peak_kind = peaks()
    # every rank codes its own stretch of the text stream
    host_np = datagen.text_stream(total, first=rank * total)
    host_in = torch.from_numpy(host_np.copy()).pin_memory()
    d_in = host_in.to(dev, non_blocking=False)

    codec = DeviceCodec(z, zb, torch, dev, torch.cuda.current_stream().cuda_stream)
    codec.alloc(nb, bb)
    ctx, L = codec.ctx, codec.L
    launches = {"n": 0}
    kern_ms = {"enc": 0.0, "dec": 0.0, "enc_n": 0, "dec_n": 0}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    for _ in range(max(1, args.warmup)):
        codec.step(level, d_in)
    if not torch.equal(codec.d_plain, d_in):
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
        a, b, arc_total, st_c, st_d = codec.step(level, d_in)
        tc += a
        td += b
        launches["n"] += st_c["launches"] + st_d["launches"]
        kern_ms["enc"] += st_c["codec_ms"]; kern_ms["enc_n"] += st_c["codec_launches"]
        kern_ms["dec"] += st_d["codec_ms"]; kern_ms["dec_n"] += st_d["codec_launches"]
    barrier()
    clk = clocks.stop() if rank == 0 else {}
    t_sum, t_c, t_d = allmax([tc + td, tc, td])
    bytes_all = total * world * args.steps
    value = bytes_all / t_sum / 1e6

    # parity against the oracle (outside the timed region): 8 blocks spread over the batch, byte-identical
    parity = None
    parity_idx = sorted(set(int(x) for x in np.linspace(0, nb - 1, 8)))
    if rank == 0:
        parity = bool(parity_blocks(codec, level, host_np, parity_idx, cores))

    # ---- end to end through the host-buffer C ABI ----
    e2e = None
    if not args.no_e2e:
        cap = codec.cap
        host_arc = torch.empty(cap, dtype=torch.uint8).pin_memory()
        host_out = torch.empty(total, dtype=torch.uint8).pin_memory()
        out_off = (C.c_uint64 * (nb + 1))()
        need = C.c_uint64(0)
        comments = (C.c_char_p * nb)(*[b"%d bytes" % bb for _ in range(nb)])
        nseg = C.c_int(0)
        segs = (zb.Segment * (nb + 8))()

        def step_host():
            t0 = time.perf_counter()
            ctx._check(L.zpaqgpu_compress_blocks(ctx._h, level, host_in.data_ptr(), codec.in_off, nb, None, comments,
                                                 host_arc.data_ptr(), cap, out_off, C.byref(need)))
            t1 = time.perf_counter()
            arc_len = int(out_off[nb])
            rc = L.zpaqgpu_decompress_archive(ctx._h, host_arc.data_ptr(), arc_len, host_out.data_ptr(), total,
                                              C.byref(need), segs, nb + 8, C.byref(nseg))
            ctx._check(rc)
            t2 = time.perf_counter()
            return t1 - t0, t2 - t1, arc_len

        e_steps = max(1, min(args.steps, args.e2e_steps))
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
        t_h = allmax([hc + hd])[0]
        e2e = {"value": round(total * world * e_steps / t_h / 1e6, 3), "unit": "MB/s",
               "h2d_bytes_per_step": int(total + arc_len), "d2h_bytes_per_step": int(arc_len + total),
               "compress_mb_s": round(total * e_steps / hc / 1e6, 3),
               "decompress_mb_s": round(total * e_steps / hd / 1e6, 3), "steps": e_steps,
               "steps_note": "host-timed (perf_counter around the C-ABI calls, copies inside); capped at --e2e-steps "
                             "so that the per-level passes fit the run"}
        del host_arc, host_out

    # ---- roofline of the dominant kernel (the per-SM integer issue roofline) ----
    ratio = arc_total / total
    f_clk = (clk.get("sm_mhz") or sm_max) * 1e6 if rank == 0 else sm_max * 1e6
    roofline = None
    if rank == 0:
        enc_ms = kern_ms["enc"] / max(1, kern_ms["enc_n"])
        dec_ms = kern_ms["dec"] / max(1, kern_ms["dec_n"])
        launch_bytes = total / max(1, st_d["waves"])
        issue_peak = N_SM * LANES * f_clk
        tj = load_traffic()

        def traffic_of(key):
            if not tj or key not in tj:
                return None
            t = tj[key]
            same = t.get("level") == level and t.get("blocks") == nb and t.get("block_bytes") == bb
            return int(t["dram_bytes_per_launch"]) if same else None

        n_isse = N_HT.get(level, 3) - 1
        mix = "true" if level >= 4 else "false"
        dec_name = os.environ.get("ZPAQGPU_DECODER", "tree")
        tree_dec = level <= 3 and dec_name != "serial"
        if tree_dec and dec_name == "tree2":
            dec_kernel = "k_decode_tree2<%d>" % n_isse
        else:
            dec_kernel = "k_decode_chain<%d,%s,%s>" % (n_isse, mix, "true" if tree_dec else "false")

        def kern(ms, name, key, st):
            ach = launch_bytes / (ms / 1e3) * W_OPS[level]
            return {"kernel": name, "kernel_ms": round(ms, 3), "achieved": round(ach / 1e12, 5),
                    "frac": round(ach / issue_peak, 5), "traffic": traffic_of(key),
                    "hbm": hbm_block(level, ratio, st["workspace_bytes_per_block"], st["paged"], bb, launch_bytes,
                                     ms / 1e3, hbm_peak)}

        dec_r = kern(dec_ms, dec_kernel, "decode", st_d)
        enc_r = kern(enc_ms, "k_encode_pipe3<%d,%s>" % (n_isse, mix), "encode", st_c)
        dom, other = (dec_r, enc_r) if dec_ms >= enc_ms else (enc_r, dec_r)
        roofline = {
            "bound": "issue", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": round(issue_peak / 1e12, 4),
            "unit": "Tops/s", "frac": dom["frac"], "traffic": dom["traffic"], "kernel_ms": dom["kernel_ms"],
            "ops_per_input_byte": W_OPS[level], "units_per_launch": int(launch_bytes),
            "peak_source": "148 SM x 128 INT32 lanes x f_clk (SURVEY 8(d)); f_clk = median SM clock sampled during the "
                           "timed region (%.0f MHz)" % (f_clk / 1e6),
            "formula": "achieved = W(L) ops/B x units_per_launch / kernel time (CUDA events around the launch); "
                       "frac = achieved / peak",
            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch on this configuration "
                              "(profiles/r02_traffic_launches.csv)" if dom["traffic"] else None,
            "hbm": dict(dom["hbm"], peak_source=peak_kind), "other_kernel": other,
        }

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        n_s = min(nb, cpu_sample_blocks(level, cores, bb))
        a, b, _, _, _ = oracle_run(level, host_np, n_s, bb, cores)
        cpu = {"value": round(n_s * bb / (a + b) / 1e6, 3), "unit": "MB/s", "cores": cores, "kind": "port",
               "sample": "first %d of %d blocks (%d KiB each), compress+decompress, CPU oracle (C restatement of the "
                         "V reference) on %d threads" % (n_s, nb, args.block_kib, cores),
               "compress_mb_s": round(n_s * bb / a / 1e6, 3), "decompress_mb_s": round(n_s * bb / b / 1e6, 3)}

    # ---- the other BASELINE.json configurations ----
    codec.close()
    del codec, d_in, host_in
    torch.cuda.empty_cache()
    per_level = None
    if not args.no_per_level:
        per_level = run_per_level(args, dict(z=z, zb=zb, torch=torch, np=np, datagen=datagen, dev=dev, rank=rank,
                                             world=world, barrier=barrier, allmax=allmax, f_clk=f_clk, cores=cores,
                                             hbm_peak=hbm_peak, text=host_np))
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(t_sum / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(args),
        "compress_mb_s": round(bytes_all / t_c / 1e6, 3), "decompress_mb_s": round(bytes_all / t_d / 1e6, 3),
        "ratio": round(ratio, 4), "byte_identical_to_oracle": parity, "parity_blocks": parity_idx,
        "e2e": e2e, "gpu_launches": launches["n"], "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
        "per_level": per_level, "stats": {"compress": st_c, "decompress": st_d},
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_per_level(args, E):
    """One pass of every other BASELINE.json configuration.  Kernel times come from CUDA events around the
    launches (zpaqgpu_last_stats); the per-call times from events around the device-pointer calls."""
    torch, np, datagen, dev = E["torch"], E["np"], E["datagen"], E["dev"]
    rank, world = E["rank"], E["world"]
    want = [x.strip() for x in args.per_level.split(",") if x.strip()]
    out = {}

    def one(key, level, data_np, nb, bb, scaling, what, warm_blocks=148):
        """data_np: this rank's input (nb * bb bytes, host).  Every rank reaches every collective: a failure on
        one rank (say, out of memory) is agreed on through the max-reduction instead of leaving the others at a
        barrier."""
        codec = None
        state = {}

        def attempt(fn):
            err = None
            try:
                fn()
            except SystemExit:
                raise
            except Exception as ex:
                err = "%s: %s" % (type(ex).__name__, ex)
            bad = E["allmax"]([1.0 if err else 0.0])[0] > 0
            return err if err else ("another rank failed" if bad else None)

        def setup():
            nonlocal codec
            codec = DeviceCodec(E["z"], E["zb"], torch, dev, torch.cuda.current_stream().cuda_stream)
            data = data_np if data_np.flags.writeable else data_np.copy()
            state["d_in"] = torch.from_numpy(data).to(dev)
            # a small pass first: module load, first-touch of the constant tables, buffer growth
            wb = min(nb, warm_blocks)
            state["wb"] = wb
            codec.alloc(wb, bb, slack=2)
            codec.step(level, state["d_in"][:wb * bb])
            codec.alloc(nb, bb, slack=2 if level == 3 else 4)
            if level <= 3 and nb > wb:
                # dense tables (-m1..-m3): the buffers for all nb blocks exist before the timed pass.  Sizing only:
                # a failure here changes nothing about what is measured, so it is noted and the pass goes on
                try:
                    codec.presize(level, state["d_in"])
                except (Exception, SystemExit) as ex:
                    state["presize_error"] = "%s: %s" % (type(ex).__name__, ex)

        def timed():
            state["res"] = codec.step(level, state["d_in"])
            if not torch.equal(codec.d_plain, state["d_in"]):
                raise RuntimeError("round trip mismatch")

        try:
            total = nb * bb
            err = attempt(setup)
            if not err:
                E["barrier"]()
                err = attempt(timed)
            if err:
                return {"what": what, "level": level, "error": err}
            t_c, t_d, arc, st_c, st_d = state["res"]
            wb = state["wb"]
            presize_error = state.get("presize_error")
            idx = sorted(set(int(x) for x in np.linspace(0, nb - 1, 8)))
            ok = bool(parity_blocks(codec, level, data_np, idx, E["cores"])) if rank == 0 else None
            m_c, m_d, k_c, k_d = E["allmax"]([t_c, t_d, st_c["codec_ms"] / 1e3, st_d["codec_ms"] / 1e3])
            all_bytes = total * world
            return {
                "what": what, "level": level, "blocks_per_gpu": nb, "block_bytes": bb, "scaling": scaling,
                "input_bytes_all_gpus": all_bytes,
                "compress_mb_s": round(all_bytes / m_c / 1e6, 2), "decompress_mb_s": round(all_bytes / m_d / 1e6, 2),
                "compress_kernel_mb_s": round(all_bytes / k_c / 1e6, 2),
                "decompress_kernel_mb_s": round(all_bytes / k_d / 1e6, 2),
                "compress_kernel_ms": round(st_c["codec_ms"], 2), "decompress_kernel_ms": round(st_d["codec_ms"], 2),
                "issue_frac": {"compress": round(issue_frac(level, total, k_c, E["f_clk"]), 5),
                               "decompress": round(issue_frac(level, total, k_d, E["f_clk"]), 5)},
                "hbm": {"compress": hbm_block(level, arc / total, st_c["workspace_bytes_per_block"], st_c["paged"], bb,
                                              total, k_c, E["hbm_peak"]),
                        "decompress": hbm_block(level, arc / total, st_d["workspace_bytes_per_block"], st_d["paged"],
                                                bb, total, k_d, E["hbm_peak"])},
                "ratio": round(arc / total, 4),
                "tables": ["paged" if st_c["paged"] else "dense", "paged" if st_d["paged"] else "dense"],
                "waves": [st_c["waves"], st_d["waves"]], "retries": [st_c["retries"], st_d["retries"]],
                "pool_mb_used": [round(st_c["pool_bytes_used"] / 1e6, 1), round(st_d["pool_bytes_used"] / 1e6, 1)],
                "warps_per_cta": [st_c["warps_per_cta"], st_d["warps_per_cta"]],
                "byte_identical_to_oracle": ok, "parity_blocks": idx, "round_trip_exact": True,
                "timing": "one pass after a %d-block warm-up pass%s; CUDA events; kernel figures = the codec kernel "
                          "alone, the others the whole device-pointer call (table clearing, SHA-1, assembly included)"
                          % (wb, " and a table-sizing pass over %d blocks of 64 bytes" % nb
                             if level <= 3 and nb > wb and not presize_error else ""),
                **({"presize_error": presize_error} if presize_error else {}),
            }
        finally:
            if codec is not None:
                codec.close()
            codec = None
            state.clear()
            torch.cuda.empty_cache()

    text = E["text"]  # this rank's 1 GiB (or --blocks x --block-kib) of the text stream
    mib = 1 << 20
    n_text_mib = len(text) // mib
    for key in want:
        try:
            if key in ("1", "4") and n_text_mib >= 1:
                out["m" + key] = one("m" + key, int(key), text[:n_text_mib * mib], n_text_mib, mib, "weak",
                                     "-m%s, %d x 1 MiB text per GPU" % (key, n_text_mib))
            elif key == "3":
                # cfg 3: the 8192 blocks of 1 MiB are split over the ranks in contiguous ranges (strong
                # scaling).  1024 distinct mixed blocks are generated and repeated to fill a rank's range.
                per = max(1, args.cfg3_blocks // world)
                distinct = min(per, 1024)
                base = datagen.mixed_stream(distinct, mib, datagen.SEED0 + 3 + 1000 * rank)
                reps = (per + distinct - 1) // distinct
                data = np.tile(base, reps)[:per * mib] if reps > 1 else base
                out["m3"] = one("m3", 3, data, per, mib, "strong",
                                "cfg 3: -m3, %d x 1 MiB mixed text/random/structured over %d GPU(s) (%d distinct "
                                "blocks per GPU, repeated)" % (per * world, world, distinct))
                del data, base
            elif key == "5":
                # cfg 4: -m5, 4 MiB blocks of text; the rank's text stream repeated to fill the blocks
                bb = 4 * mib
                nbk = max(1, args.cfg4_blocks)
                have = (len(text) // bb) * bb
                if have == 0:
                    continue
                reps = (nbk * bb + have - 1) // have
                data = np.tile(text[:have], reps)[:nbk * bb] if reps > 1 else text[:nbk * bb]
                out["m5"] = one("m5", 5, np.ascontiguousarray(data), nbk, bb, "weak",
                                "cfg 4: -m5, %d x 4 MiB text per GPU (%d distinct blocks, repeated); decompression "
                                "is the quoted figure" % (nbk, have // bb), warm_blocks=16)
                del data
            elif key == "jidac" and rank == 0:
                out["jidac_add"] = run_jidac(args, E)
            elif key == "generic" and rank == 0:
                out["generic"] = run_generic(args, E)
        except SystemExit:
            raise
        except Exception as ex:  # rank-0-only sections (jidac, generic): reported, never silently dropped
            if key not in ("jidac", "generic"):
                raise
            out[{"jidac": "jidac_add", "generic": "generic"}[key]] = {"error": "%s: %s" % (type(ex).__name__, ex)}
            torch.cuda.empty_cache()
    E["barrier"]()
    return out


GENERIC_HEADERS = {
    # tests/test_oracle_kats.py CUSTOM_HEADERS: five components of five different types (nothing for the lanes
    # to share) and twenty components with eight ICMs and four ISSEs (lanes of one type run together)
    "icm_match_mix2_sse": [3, 10, 0, 0, 5, 3, 14, 4, 12, 14, 6, 10, 0, 1, 20, 255, 8, 12, 2, 9, 10, 3, 32, 100, 0,
                           104, 17, 28, 59, 112, 25, 60, 25, 59, 112, 25, 65, 112, 25, 112, 56, 0],
    "twenty": ([5, 10, 0, 0, 20] + [3, 10, 3, 11, 3, 12, 3, 10, 3, 11, 3, 12, 3, 10, 3, 11]
               + [8, 11, 0, 8, 11, 8, 8, 10, 9, 8, 10, 3] + [4, 12, 14] + [2, 12, 20, 2, 14, 255] + [5, 10, 11, 128]
               + [9, 8, 15, 16, 64] + [7, 8, 0, 17, 20, 255] + [6, 10, 16, 17, 24, 255] + [9, 10, 18, 32, 255] + [0]
               + [96, 4, 28] + [59, 112, 25, 10] * 6 + [59, 112, 25] * 8 + [60, 25] * 3 + [59, 112, 25] * 3 + [56, 0]),
}


def run_generic(args, E):
    """headers outside the -m1..-m5 shape through the generic path: the warp kernel (component
    per lane, warp-uniform ZPAQL) beside the one-lane kernel it replaces (ZPAQGPU_GENERIC=lane0), same blocks,
    kernel time from CUDA events."""
    z = E["z"]
    import oracle_binding as ob
    nb, bb = 592, 32768
    text = E["text"]
    blocks = [text[k * bb:(k + 1) * bb].tobytes() for k in range(nb)]
    out = {"what": "custom headers through the generic kernels, %d x %d KiB text blocks, host buffers; warp = "
                   "kernels_genwarp.cu, lane0 = the one-lane kernel" % (nb, bb // 1024), "blocks": nb, "block_bytes": bb}
    for hname, hbytes in GENERIC_HEADERS.items():
        hdr = bytes(hbytes)
        res, arcs = {}, {}
        for name, env in (("warp", None), ("lane0", "lane0")):
            if env:
                os.environ["ZPAQGPU_GENERIC"] = env
            else:
                os.environ.pop("ZPAQGPU_GENERIC", None)
            ctx = z.Context(E["dev"].index)
            try:
                ctx.compress_blocks(0, blocks[:8], header=hdr)
                got = ctx.compress_blocks(0, blocks, header=hdr)
                st_c = ctx.stats()
                plain, segs, status = ctx.decompress_archive(b"".join(got))
                st_d = ctx.stats()
                if status != 0 or plain != b"".join(blocks):
                    raise SystemExit("generic %s/%s: round trip mismatch" % (hname, name))
                arcs[name] = got
                res[name] = {"compress_kernel_ms": round(st_c["codec_ms"], 2), "decompress_kernel_ms": round(st_d["codec_ms"], 2),
                             "compress_kernel_mb_s": round(nb * bb / st_c["codec_ms"] / 1e3, 2),
                             "decompress_kernel_mb_s": round(nb * bb / st_d["codec_ms"] / 1e3, 2)}
            finally:
                ctx.close()
        os.environ.pop("ZPAQGPU_GENERIC", None)
        res["speedup_over_lane0"] = {
            "compress": round(res["lane0"]["compress_kernel_ms"] / res["warp"]["compress_kernel_ms"], 2),
            "decompress": round(res["lane0"]["decompress_kernel_ms"] / res["warp"]["decompress_kernel_ms"], 2)}
        idx = [0, 77, 300, 591]
        res["byte_identical_to_oracle"] = all(
            arcs["warp"][k] == arcs["lane0"][k] == ob.compress_block(0, blocks[k], "", "", header=hdr) for k in idx)
        res["parity_blocks"] = idx
        res["components"] = hbytes[4]
        out[hname] = res
    return out


def run_jidac(args, E):
    """cfg 5: jidac add of the synthetic tree through zpaqgpu_jidac_add with HOST buffers (rank 0's GPU)."""
    np, datagen = E["np"], E["datagen"]
    import oracle_binding as ob
    z, zb = E["z"], E["zb"]
    DATE = 20260101120000
    names, files = datagen.file_tree(args.jidac_files)
    total = sum(map(len, files))
    kw = dict(level=1, fragment=6, dedup=True, block_bytes=1 << 20)
    ctx = z.Context(E["dev"].index)
    try:
        ctx.jidac_add(names[:4], files[:4], DATE, **kw)
        src = np.frombuffer(b"".join(files), dtype=np.uint8)
        off = np.zeros(len(files) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(f) for f in files])
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        opts = zb.JidacOpts(DATE, kw["level"], kw["fragment"], 1, 0, kw["block_bytes"])
        outb = np.empty(total + total // 4 + 4096 * (len(files) + 4), dtype=np.uint8)
        ln, need = C.c_uint64(0), C.c_uint64(0)
        for rep in range(2):  # the first call grows the device buffers (slow once peers are mapped under NCCL)
            t0 = time.perf_counter()
            rc = zb.lib().zpaqgpu_jidac_add(ctx._h, C.byref(opts), arr, src.ctypes.data, off.ctypes.data, len(files),
                                            outb.ctypes.data, outb.nbytes, C.byref(ln), C.byref(need))
            t1 = time.perf_counter()
            ctx._check(rc)
        st = ctx.jidac_stats()
        k = min(100, len(files))
        want = ob.jidac_add(names[:k], files[:k], DATE, **kw)
        got = ctx.jidac_add(names[:k], files[:k], DATE, **kw)
        return {"what": "cfg 5: jidac add, %d files (1 KiB..1 MiB, 30 %% duplicates), fragment 6, dedup, -m1, 1 MiB d "
                        "blocks, host buffers in and out, one GPU" % len(files),
                "input_bytes": total, "add_mb_s": round(total / (t1 - t0) / 1e6, 2), "archive_bytes": int(ln.value),
                "stages_ms": {s: round(st[s], 2) for s in ("h2d_ms", "fragment_ms", "sha1_ms", "dedup_ms", "gather_ms",
                                                           "codec_ms", "pack_ms", "d2h_ms")},
                "n_fragments": st["n_fragments"], "n_stored": st["n_stored"], "n_dblocks": st["n_dblocks"],
                "stored_bytes": st["stored_bytes"], "launches": st["launches"],
                "byte_identical_to_oracle_on_subtree": got == want, "subtree_files": k}
    finally:
        ctx.close()


if __name__ == "__main__":
    sys.exit(main())
