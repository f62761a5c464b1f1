"""bench.py -- FAME hot-path benchmark (driver contract: ONE JSON line on stdout from rank 0).

BASELINE.json's metric has two halves, "FAME train patients/sec & 512-tok note chunks/sec at 1/2/4/8 B200":

  headline   : the full FAME training step (BASELINE configs[3]): 32 patients per GPU, L = 542 lab tokens, three tasks,
               forward + BCE / LEDDI loss + hand-written backward + clip + AdamW with the reference's dropout, data
               parallel with the EDDI-statistic exchange and the gradient reduction -- the half that has collectives,
               so the driver's 1 -> 8 scaling run measures something.  metric / value / e2e / roofline / cpu_baseline
               of the line describe THIS step.
  note_encoder: the second half (BASELINE configs[1]): BioClinicalBERT note-chunk encoder forward, 256 chunks x 512
               tokens per GPU + chunk -> patient pool, as a complete sub-object with its own value / e2e / roofline /
               cpu_baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 4|2|3|5]

  --config 4 (default) the line above           --config 2  the note encoder alone as the headline
  --config 3  BEHRT towers forward + backward at 1024 patients on one GPU (BASELINE configs[2])
  --config 5  46 k-patient evaluation sweep over N GPUs (BASELINE configs[4])

  value : whole-job throughput with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e   : the same through the public API from pinned HOST buffers (H2D of the step's inputs and D2H of its result
          inside the timed region)
  --impl reference : the reference's CPU path for the same step on the box's host threads (rank 0 only): the oracle
          port of train_step under torch.autograd with torch's own clip_grad_norm_ / AdamW (the reference script cannot
          travel: HF hub + MIMIC CSVs), and transformers.BertModel called one chunk at a time exactly as
          10_FAME.py:140,157-169 does for the note encoder.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNKS, SEQ, CHUNKS_PER_PATIENT = 256, 512, 4
NOTE_WORKLOAD = ("note_encoder_fwd_bert_base_256chunks_x_512tok_bf16 + chunk->patient mean pool (BASELINE configs[1]); the last "
                 "layer runs for the CLS rows only (the reference reads last_hidden_state[:, 0, :], 10_FAME.py:141)")
NOTE_METRIC = "512-tok note chunks/sec (BioClinicalBERT note-chunk encoder forward)"
TRAIN_WORKLOAD = ("fame_train_step_32patients_per_gpu_L542_3tasks: forward + BCE/LEDDI loss + backward + clip + AdamW, "
                  "text embeddings precomputed as in 10_FAME.py:729-731 (BASELINE configs[3])")
TRAIN_METRIC = "FAME train patients/sec (full training step)"
FLOP_PER_TOKEN = 188_743_680           # 12*(2*(4*768^2 + 2*768*3072) + 4*512*768), all 12 layers, all rows
LAB_FLOP_PER_TOKEN = 25_350_144        # lab tower forward at L = 542
DEMO_FLOP_PER_PATIENT = 169_900_000    # demographic tower forward (Q / K projections included)
WSEED = 7
TRAIN_B, TRAIN_L = 32, 542
KEYS9 = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids",
         "lab_features", "text", "labels")


def note_flop_executed(chunks=CHUNKS, seq=SEQ, layers=12):
    """FLOPs the note encoder EXECUTES per step with the CLS-only last layer: 11 full layers, then K / V projections of
    all rows and everything else for `chunks` rows."""
    H, F = 768, 3072
    full = 2 * (4 * H * H + 2 * H * F) + 4 * seq * H                   # per token, one layer
    last = chunks * seq * 2 * (2 * H * H) + chunks * (2 * (2 * H * H + 2 * H * F) + 4 * seq * H)
    return (layers - 1) * full * chunks * seq + last


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions only (`with sampler.region():` around every timed
    loop): NVML from a background thread every ~3 ms while a region is open.  Falls back to one nvidia-smi loop over
    the whole run when NVML cannot be loaded."""
    NAMES = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self.active, self.stop_flag = index, [], 0, set(), False, False
        self.nvml = self.handle = self.thread = self.smi = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None
            self._start_smi()

    def _loop(self):
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            if self.active:
                try:
                    self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                    r = int(reasons(self.handle))
                    for name, bit in self.NAMES:
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.003)
            else:
                time.sleep(0.0005)

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.lines = []
        try:
            self.smi = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                         "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.smi.stdout], daemon=True).start()
        except Exception:
            self.smi = None

    def region(self):
        import contextlib

        @contextlib.contextmanager
        def cm():
            self.active = True
            try:
                yield
            finally:
                self.active = False
        return cm()

    def stop(self):
        self.stop_flag = True
        if self.nvml is None:
            if self.smi is None:
                return None
            time.sleep(0.12)
            self.smi.terminate()
            for l in self.lines:
                f = [x.strip() for x in l.split(",")]
                try:
                    self.sm.append(float(f[0])); self.mx = max(self.mx, float(f[1]))
                except (ValueError, IndexError):
                    continue
                for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
        if not self.sm:
            return None
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml, inside the timed regions only" if self.nvml is not None else "nvidia-smi, whole run"}


class _NoSampler:
    def region(self):
        import contextlib
        return contextlib.nullcontext()

    def stop(self):
        return None


def host_threads():
    """Host threads available to this process.  torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the
    reference arm on one core: the CPU legs set the thread count explicitly instead."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ====================================================================================================== CPU legs
def _hf_note_encoder():
    """transformers.BertModel -- the module the reference wraps (10_FAME.py:133-142, 726-728) -- with the bench's
    synthetic weights.  Returns (callable(ids, mask) -> CLS rows, description)."""
    import torch

    from fairmultimodal_b200 import synth
    sd = {k[len("BioBert."):]: torch.from_numpy(v) for k, v in
          synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), WSEED).items()}
    try:
        from transformers import BertConfig, BertModel
        bert = BertModel(BertConfig(vocab_size=synth.VOCAB))
        bert.load_state_dict(sd, strict=True)
        bert.eval()
        return (lambda ids, mask: bert(input_ids=ids, attention_mask=mask).last_hidden_state[:, 0, :],
                "transformers.BertModel (stock, fp32 eager; what 10_FAME.py:140 calls)", "reference")
    except Exception as e:                                   # transformers missing on the box: the oracle restatement
        from oracle import fame_oracle as O
        full = {"BioBert." + k: v for k, v in sd.items()}
        return (lambda ids, mask: O.note_cls(full, ids, mask),
                f"oracle port of BioClinicalBERT_FT.forward (transformers unavailable: {type(e).__name__})", "port")


def cpu_note_chunks_per_s(n_chunks, threads=None, warm=1):
    """The reference's CPU path for the note encoder: one chunk per call (10_FAME.py:157-169), fp32, eager.
    Returns (chunks/s, threads, seconds, description, kind)."""
    import torch

    from fairmultimodal_b200 import synth
    torch.set_num_threads(threads or host_threads())
    fn, desc, kind = _hf_note_encoder()
    co = synth.make_cohort(max(1, (n_chunks + 3) // 4), lab_tokens=4, chunks="fixed4", seq_len=SEQ, seed=1234)
    ids, mask = torch.from_numpy(co["input_ids"]), torch.from_numpy(co["attention_mask"])
    with torch.no_grad():
        for j in range(warm):
            fn(ids[j:j + 1], mask[j:j + 1])                  # warm the thread pool / allocator
        t0 = time.perf_counter()
        for j in range(n_chunks):
            fn(ids[j:j + 1], mask[j:j + 1])
        dt = time.perf_counter() - t0
    return n_chunks / dt, torch.get_num_threads(), dt, desc, kind


class CpuTrainStep:
    """The reference's CPU path for the training step (10_FAME.py:401-449 on the CPU device): forward of the three
    modules + BCE / LEDDI loss + autograd backward + clip_grad_norm_(1.0) + AdamW, fp32 eager, 32 patients, L = 542 --
    the oracle's forward / loss under torch.autograd, torch's own clip and AdamW.  Dropout off (the oracle restates
    the eval-mode arithmetic), which only makes the CPU side cheaper."""

    def __init__(self, threads=None):
        import numpy as np
        import torch

        from fairmultimodal_b200 import synth
        from oracle import fame_oracle as O
        torch.set_num_threads(threads or host_threads())
        self.torch, self.O = torch, O
        shapes = synth.fame_shapes(lab_tokens=TRAIN_L)
        self.sd = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in synth.synth_state_dict(shapes, 4).items()}
        co = synth.make_cohort(TRAIN_B, lab_tokens=TRAIN_L, chunks=0, with_tokens=False, seed=77)
        co["text"] = np.random.default_rng(0).standard_normal((TRAIN_B, 768)).astype(np.float32)
        self.batch = [torch.from_numpy(co[k]) for k in KEYS9]
        self.pw = torch.from_numpy(synth.pos_weight(co["labels"]))
        self.params = list(self.sd.values())
        self.opt = torch.optim.AdamW(self.params, lr=1e-5, weight_decay=0.01)

    def step(self):
        torch, O, b = self.torch, self.O, self.batch
        self.opt.zero_grad()
        o = O.fame_forward(self.sd, b, (0.33, 0.33, 0.33))
        total, _, _ = O.fame_loss(o["fused_logits"], b[8], (b[2], b[4], b[5]), self.sd["sig_weights"], self.pw, 0.8, 0.01)
        total.backward()
        torch.nn.utils.clip_grad_norm_([p for p in self.params if p.grad is not None], 1.0)
        self.opt.step()

    def time(self, steps, warm=1):
        for _ in range(warm):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        dt = time.perf_counter() - t0
        return steps * TRAIN_B / dt, self.torch.get_num_threads(), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == 2:
        per_step = args.ref_chunks_per_step
        v, threads, dt, desc, kind = cpu_note_chunks_per_s(per_step * args.steps, warm=max(1, min(args.warmup, 3)))
        line = {"impl": "reference", "metric": NOTE_METRIC, "value": v, "unit": "chunks/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": NOTE_WORKLOAD, "sample": f"{per_step} chunk(s) per step, one chunk per call"},
                "cpu_baseline": {"value": v, "unit": "chunks/s", "cores": threads, "kind": kind,
                                 "sample": f"{per_step * args.steps} chunks x 512 tokens, {desc}, one chunk per call"},
                "e2e": {"value": v, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    if args.config != 4:
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU reference arm covers --config 4 and 2 (asked: {args.config}); "
                          "configs 3 / 5 are hours of CPU work"}), flush=True)
        return
    cpu = CpuTrainStep()
    v, threads, dt = cpu.time(args.steps, warm=max(1, min(args.warmup, 2)))
    sample = (f"{args.steps} full step(s) of 32 patients, L = 542 ({dt:.1f} s) after {max(1, min(args.warmup, 2))} warm-up step(s): "
              "oracle forward + loss under torch.autograd, clip_grad_norm_, torch AdamW, fp32, dropout off")
    n_note = max(8, min(40, 2 * args.steps))
    nv, _, ndt, desc, kind = cpu_note_chunks_per_s(n_note)
    line = {
        "impl": "reference", "metric": TRAIN_METRIC, "value": v, "unit": "patients/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": TRAIN_WORKLOAD, "patients_per_step": TRAIN_B, "lab_tokens": TRAIN_L,
                   "note": "CPU arm: one process, one 32-patient batch per step at every N (no data parallelism on the host)"},
        "cpu_baseline": {"value": v, "unit": "patients/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "patients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note_encoder": {"metric": NOTE_METRIC, "value": nv, "unit": "chunks/s",
                         "cpu_baseline": {"value": nv, "unit": "chunks/s", "cores": threads, "kind": kind,
                                          "sample": f"{n_note} chunks x 512 tokens ({ndt:.1f} s), {desc}, one chunk per call "
                                                    "as 10_FAME.py:157-169"}},
    }
    print(json.dumps(line), flush=True)


# ====================================================================================================== helpers
def _kernel_table(trace, steps):
    """Per C-ABI op (GEMM launches split into the tcgen05 kernel and the <= 32-row weight-streaming kernel): device
    time, launches and algorithmic work from the CUDA-event pairs around every launch of the trace pass."""
    by = {}
    for name, tag, a, b, work in trace:
        key = name
        if name in ("fame_gemm_bias_act", "fame_gemm_ex"):
            key = "gemm_bf16_tcgen05_kernel" if tag.startswith("tc") else "skinny_gemm_kernel"
        d = by.setdefault(key, [0.0, 0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += work; d[2] += 1
    tot = sum(v[0] for v in by.values()) or 1.0
    table = {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[2] / steps, "share": v[0] / tot}
             for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])}
    # the tcgen05 GEMM by role (forward layers / data gradients / weight gradients / the attention backward's batched
    # products): time and achieved TFLOP/s of each group
    roles = {}
    for name, tag, a, b, work in trace:
        if name in ("fame_gemm_bias_act", "fame_gemm_ex") and tag.startswith("tc"):
            parts = tag.split(":")
            role = parts[1] if len(parts) == 3 and parts[1] else "fwd"
            d = roles.setdefault(role, [0.0, 0.0, 0])
            d[0] += a.elapsed_time(b); d[1] += work; d[2] += 1
    if roles:
        table["gemm_bf16_tcgen05_kernel"]["by_role"] = {
            r: {"ms_per_step": v[0] / steps, "launches_per_step": v[2] / steps, "tflops": v[1] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0}
            for r, v in sorted(roles.items(), key=lambda kv: -kv[1][0])}
    return by, table


def _traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of the same command
    (profiles/roofline_traffic.json, written by scripts/summarize_profiles.py; keys '<workload>:<kernel>')."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        t = json.load(open(path))
        return t.get(kernel), t.get("_source_" + kernel)
    return None, None


def _gemm_roofline(by, pk, workload_key):
    g = by.get("gemm_bf16_tcgen05_kernel")
    if not g or g[0] <= 0:
        return None
    tf = g[1] / (g[0] * 1e-3) / 1e12
    traffic, src = _traffic(workload_key)
    return {"kernel": "gemm_bf16_tcgen05_kernel", "bound": "tensor", "achieved": tf, "peak": pk["tf_sust"], "unit": "TFLOP/s",
            "frac": tf / pk["tf_sust"], "traffic": traffic, "traffic_source": src,
            "peak_source": f"{pk['src']} (sustained cuBLAS bf16: the kernel is timed inside a long step)",
            "launches": g[2], "avg_launch_ms": g[0] / g[2],
            "timing": "CUDA events around every launch on its launching stream, separate trace pass (not the value region)"}


def _max_over_ranks(vals, world, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


# ====================================================================================================== note encoder
def bench_note_encoder(args, world, rank, dev, barrier, pk, sampler):
    import torch

    from fairmultimodal_b200 import modules, ops, synth
    sd = {k: torch.from_numpy(v) for k, v in
          synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), WSEED).items()}
    model = modules.BioClinicalBERT_FT.from_state_dict(sd).to(dev)
    del sd
    patients = CHUNKS // CHUNKS_PER_PATIENT
    n_batches = 2                                               # rotate input batches between steps
    co = synth.make_cohort(patients * n_batches, lab_tokens=4, chunks="fixed4", seq_len=SEQ, seed=1234 + rank)
    ids_h = torch.from_numpy(co["input_ids"]).view(n_batches, CHUNKS, SEQ).pin_memory()
    mask_h = torch.from_numpy(co["attention_mask"]).view(n_batches, CHUNKS, SEQ).pin_memory()
    offs = torch.arange(0, CHUNKS + 1, CHUNKS_PER_PATIENT, dtype=torch.int32, device=dev)
    ids_d, mask_d = ids_h.to(dev), mask_h.to(dev)
    out_h = torch.empty((patients, 768), dtype=torch.float32).pin_memory()

    def step_resident(i):
        return modules.pool_chunks(model.encode_cls(ids_d[i % n_batches], mask_d[i % n_batches]), offs)

    def step_e2e(i):
        ids = ids_h[i % n_batches].to(dev, non_blocking=True)
        mask = mask_h[i % n_batches].to(dev, non_blocking=True)
        out_h.copy_(modules.pool_chunks(model.encode_cls(ids, mask), offs), non_blocking=True)

    warm = max(args.warmup, 3)
    for i in range(warm):
        step_resident(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = ops.LAUNCHES
    with sampler.region():
        e0.record()
        for i in range(args.steps):
            step_resident(i)
        e1.record()
        barrier()
    launches = ops.LAUNCHES - l0
    ms = e0.elapsed_time(e1)
    for i in range(2):
        step_e2e(i)
    barrier()
    with sampler.region():
        e0.record()
        for i in range(args.steps):
            step_e2e(i)
        e1.record()
        barrier()
    ms_e2e = e0.elapsed_time(e1)
    # third pass, outside both timed regions: CUDA events around every launch (roofline of the dominant kernel)
    tsteps = min(args.steps, 5)
    ops.start_trace()
    for i in range(tsteps):
        step_resident(i)
    torch.cuda.synchronize()
    by, table = _kernel_table(ops.stop_trace(), tsteps)
    ms, ms_e2e = _max_over_ranks([ms, ms_e2e], world, dev)
    del model, ids_d, mask_d
    torch.cuda.empty_cache()
    value = world * CHUNKS * args.steps / (ms * 1e-3)
    flop = note_flop_executed()
    return {
        "metric": NOTE_METRIC, "value": value, "unit": "chunks/s", "ms_per_step": ms / args.steps, "steps": args.steps,
        "warmup": warm, "dtype": "bf16", "scaling": "weak",
        "config": {"workload": NOTE_WORKLOAD, "chunks_per_gpu_per_step": CHUNKS, "seq_len": SEQ,
                   "patients_per_gpu_per_step": patients, "parallelism": f"dp{world} (chunks sharded, no collective)",
                   "weights": "random-init BERT-base, vocab 28996 (no checkpoint reachable)",
                   "l2": "per-step working set (216 MB bf16 weights + >1 GB activations) exceeds the 126 MB L2; "
                         "input batches rotate between steps"},
        "flop_per_step_executed": flop, "flop_per_step_all_rows_all_layers": CHUNKS * SEQ * FLOP_PER_TOKEN,
        "tflops_executed": value / world / CHUNKS * flop / 1e12 * world,
        "tensor_frac_of_sustained_peak": value / world / CHUNKS * flop / 1e12 / pk["tf_sust"],
        "e2e": {"value": world * CHUNKS * args.steps / (ms_e2e * 1e-3), "unit": "chunks/s",
                "h2d_bytes_per_step": int(ids_h[0].numel() * 8 + mask_h[0].numel() * 8),
                "d2h_bytes_per_step": int(out_h.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "roofline": _gemm_roofline(by, pk, "note_encoder:gemm_bf16_tcgen05_kernel"), "kernels": table,
    }


# ====================================================================================================== train step
def _fame_model(dev, L, no_dropout):
    import torch

    from fairmultimodal_b200 import modules
    torch.manual_seed(0)
    demo = modules.BEHRTModel_Demo(5, 2, 5, 5)
    lab = modules.BEHRTModel_Lab(L)
    model = modules.MultimodalTransformer_EDDI_Sigmoid(768, demo, lab, dev).to(dev)
    if no_dropout:
        modules.set_dropout(model, 0.0)
    return model.train()


def _train_batches(B, L, n_batches, rank, dev):
    import numpy as np
    import torch

    from fairmultimodal_b200 import synth
    co = synth.make_cohort(B * n_batches, lab_tokens=L, chunks=0, with_tokens=False, seed=77 + rank)
    co["text"] = np.random.default_rng(rank).standard_normal((B * n_batches, 768)).astype(np.float32)
    host = [[torch.from_numpy(co[k][i * B:(i + 1) * B]).pin_memory() for k in KEYS9] for i in range(n_batches)]
    devb = [[x.to(dev) for x in b] for b in host]
    pw = torch.from_numpy(synth.pos_weight(co["labels"])).to(dev)
    return host, devb, pw


def bench_train(args, world, rank, dev, barrier, pk, sampler):
    """FAME training step (forward + BCE/LEDDI loss + backward + clip + AdamW), 32 patients per GPU, L = 542."""
    import torch
    import torch.distributed as dist

    from fairmultimodal_b200 import ops, train
    model = _fame_model(dev, TRAIN_L, args.no_dropout)
    n_batches = 4
    host, devb, pw = _train_batches(TRAIN_B, TRAIN_L, n_batches, rank, dev)
    hp = dict(lr=1e-5, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)
    w = (0.33, 0.33, 0.33)
    group = dist.group.WORLD if world > 1 else None
    out_h = torch.empty(4, dtype=torch.float32).pin_memory()

    def step(i, from_host, use_graph=None):
        b = [x.to(dev, non_blocking=True) for x in host[i % n_batches]] if from_host else devb[i % n_batches]
        loss = train.optimisation_step(model, b, pw, 0.8, 0.01, w, hp, group=group, use_graph=use_graph)
        if from_host:
            out_h.copy_(loss, non_blocking=True)

    warm = max(args.warmup, 3) + 1                               # first call eager, second captures, then replays
    res = {}
    for name, from_host in (("resident", False), ("e2e", True)):
        for i in range(warm if name == "resident" else 2):
            step(i, from_host)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with sampler.region():
            e0.record()
            for i in range(args.steps):
                step(i, from_host)
            e1.record()
            barrier()
        res[name] = e0.elapsed_time(e1)
    # ---- variant 4b: the note encoder inside the loop -- every step first encodes the step's 4 chunks per
    # patient (128 chunks x 512 tokens, no gradient) and pools them into the text embedding the step consumes; the
    # reference precomputes these once (10_FAME.py:729-731), so 4a above is the reference-faithful number
    in_loop = None
    if not args.skip_note_encoder:
        from fairmultimodal_b200 import modules, synth
        sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), WSEED).items()}
        enc = modules.BioClinicalBERT_FT.from_state_dict(sd).to(dev)
        del sd
        co = synth.make_cohort(TRAIN_B * n_batches, lab_tokens=4, chunks="fixed4", seq_len=SEQ, seed=4321 + rank)
        cpb = TRAIN_B * CHUNKS_PER_PATIENT
        ids_d = torch.from_numpy(co["input_ids"]).view(n_batches, cpb, SEQ).to(dev)
        mask_d = torch.from_numpy(co["attention_mask"]).view(n_batches, cpb, SEQ).to(dev)
        offs = torch.arange(0, cpb + 1, CHUNKS_PER_PATIENT, dtype=torch.int32, device=dev)

        def step_in_loop(i):
            k = i % n_batches
            text = modules.pool_chunks(enc.encode_cls(ids_d[k], mask_d[k]), offs)
            b = list(devb[k])
            b[7] = text
            train.optimisation_step(model, b, pw, 0.8, 0.01, w, hp, group=group)

        for i in range(3):
            step_in_loop(i)
        barrier()
        with sampler.region():
            e0.record()
            for i in range(args.steps):
                step_in_loop(i)
            e1.record()
            barrier()
        ms_loop = _max_over_ranks([e0.elapsed_time(e1)], world, dev)[0]
        in_loop = {"value": world * TRAIN_B * args.steps / (ms_loop * 1e-3), "unit": "patients/s", "ms_per_step": ms_loop / args.steps,
                   "workload": f"variant 4b: {cpb} note chunks x {SEQ} tokens encoded and pooled (no gradient) + the training "
                               "step, per step and GPU"}
        del enc, ids_d, mask_d
        torch.cuda.empty_cache()
    # trace pass (eager, outside the timed regions): CUDA events around every launch, on the stream it is launched on
    tsteps = 3
    l0 = ops.LAUNCHES
    ops.start_trace()
    for i in range(tsteps):
        step(i, False, use_graph=False)
    torch.cuda.synchronize()
    by, table = _kernel_table(ops.stop_trace(), tsteps)
    launches_per_step = (ops.LAUNCHES - l0) // tsteps
    res["resident"], res["e2e"] = _max_over_ranks([res["resident"], res["e2e"]], world, dev)
    st = train.get_state(model)
    n_params = int(st.n)
    info_dp = train.describe_parallel(model, group) if hasattr(train, "describe_parallel") else None
    train.release_graphs(model)            # before the process group goes away (captured NCCL kernels)
    ms_step = res["resident"] / args.steps
    flop = 3 * TRAIN_B * TRAIN_L * LAB_FLOP_PER_TOKEN                       # lab tower forward + backward (1.32 TFLOP)
    del model
    torch.cuda.empty_cache()
    return {
        "metric": TRAIN_METRIC, "value": world * TRAIN_B * args.steps / (res["resident"] * 1e-3), "unit": "patients/s",
        "ms_per_step": ms_step, "steps": args.steps, "warmup": warm,
        "e2e": {"value": world * TRAIN_B * args.steps / (res["e2e"] * 1e-3), "unit": "patients/s",
                "ms_per_step": res["e2e"] / args.steps,
                "h2d_bytes_per_step": int(sum(x.numel() * x.element_size() for x in host[0])), "d2h_bytes_per_step": 16},
        "config": {"workload": TRAIN_WORKLOAD, "patients_per_gpu_per_step": TRAIN_B, "lab_tokens": TRAIN_L,
                   "global_batch": world * TRAIN_B, "params": n_params,
                   "parallelism": f"dp{world}" + ("" if world == 1 else " (" + (info_dp or "statistic + gradient all-reduce") + ")"),
                   "cuda_graph": bool(train.USE_CUDA_GRAPH and (group is None or train._GRAPH_WITH_COLLECTIVES)),
                   "dropout": "off (parity configuration)" if args.no_dropout else
                              "0.1 at every site of the reference's train() mode (masks from an in-kernel counter hash)",
                   "precision": "bf16 GEMM operands, fp32 accumulation, fp32 master weights / optimizer state",
                   "l2": "per-step working set (196 MB bf16 weights + 392 MB fp32 masters + 1.2 GB optimizer state + "
                         "activations) exceeds the 126 MB L2; four input batches rotate between steps"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": _gemm_roofline(by, pk, "train_step:gemm_bf16_tcgen05_kernel"),
        "step_tensor": {"flop_per_step": flop, "achieved_tflops": flop / (ms_step * 1e-3) / 1e12,
                        "frac_of_sustained_peak": flop / (ms_step * 1e-3) / 1e12 / pk["tf_sust"],
                        "note": "lab-tower encoder FLOPs (SURVEY 8d: 25.35 MFLOP/token forward, x3 with backward) over the "
                                "WHOLE step time (both towers, head, loss, clip, AdamW); the demographic tower at 32 rows "
                                "is weight streaming, not tensor-bound work, and is not counted"},
        "kernels": table,
        "encoder_in_loop": in_loop,
    }


# ====================================================================================================== config 3
def bench_config3(args, world, rank, dev, barrier, pk, sampler):
    """BASELINE configs[2]: BEHRT structured encoder + demographics, 1024 patients, forward + backward, one GPU."""
    import torch

    from fairmultimodal_b200 import ops, train
    B, L = args.c3_patients, TRAIN_L
    model = _fame_model(dev, L, args.no_dropout)
    host, devb, pw = _train_batches(B, L, 2, rank, dev)
    w = (0.33, 0.33, 0.33)
    out_h = torch.empty(4, dtype=torch.float32).pin_memory()

    def step(i, from_host):
        b = [x.to(dev, non_blocking=True) for x in host[i % 2]] if from_host else devb[i % 2]
        loss, _ = train.forward_backward(model, b, pw, 0.8, 0.01, w)
        if from_host:
            out_h.copy_(loss, non_blocking=True)

    warm = max(args.warmup, 3)
    res = {}
    for name, from_host in (("resident", False), ("e2e", True)):
        for i in range(warm if name == "resident" else 2):
            step(i, from_host)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with sampler.region():
            e0.record()
            for i in range(args.steps):
                step(i, from_host)
            e1.record()
            barrier()
        res[name] = e0.elapsed_time(e1)
    l0 = ops.LAUNCHES
    ops.start_trace()
    for i in range(2):
        step(i, False)
    torch.cuda.synchronize()
    by, table = _kernel_table(ops.stop_trace(), 2)
    launches_per_step = (ops.LAUNCHES - l0) // 2
    ms_step = res["resident"] / args.steps
    flop = 3 * B * L * LAB_FLOP_PER_TOKEN + 3 * B * DEMO_FLOP_PER_PATIENT
    return {
        "metric": "BEHRT structured + demographic encoders forward+backward patients/sec", "value": B * args.steps / (res["resident"] * 1e-3),
        "unit": "patients/s", "n_gpus": 1, "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"behrt_lab(L={L}) + behrt_demo + fusion head + loss, {B} patients, forward + backward, no optimizer "
                               "(BASELINE configs[2])", "patients": B, "lab_tokens": L,
                   "dropout": "off" if args.no_dropout else "0.1 (reference train() mode)",
                   "l2": "activations of one step (>10 GB) exceed the 126 MB L2; two input batches alternate"},
        "e2e": {"value": B * args.steps / (res["e2e"] * 1e-3), "unit": "patients/s", "ms_per_step": res["e2e"] / args.steps,
                "h2d_bytes_per_step": int(sum(x.numel() * x.element_size() for x in host[0])), "d2h_bytes_per_step": 16},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": _gemm_roofline(by, pk, "config3:gemm_bf16_tcgen05_kernel"),
        "step_tensor": {"flop_per_step": flop, "achieved_tflops": flop / (ms_step * 1e-3) / 1e12,
                        "frac_of_sustained_peak": flop / (ms_step * 1e-3) / 1e12 / pk["tf_sust"]},
        "kernels": table,
    }


# ====================================================================================================== config 5
def bench_config5(args, world, rank, dev, barrier, pk, sampler):
    """BASELINE configs[4]: 46 k patients x U{1..16} chunks evaluation sweep over N GPUs."""
    import torch.distributed as dist

    from fairmultimodal_b200 import ops, sweep
    group = dist.group.WORLD if world > 1 else None
    sw = sweep.EvalSweep(args.c5_patients, world, rank, dev, group)
    sw.warm()
    sw.make_resident()
    out = {}
    l0 = ops.LAUNCHES
    for name, resident in (("resident", True), ("e2e", False)):
        barrier()
        with sampler.region():
            r, ms = sw.run(resident=resident)
            barrier()
        keys = sorted(ms)
        out[name] = dict(zip(keys, _max_over_ranks([ms[k] for k in keys], world, dev)))
        out["result"] = r
    launches = ops.LAUNCHES - l0
    P, C = sw.P, sw.chunks_total
    tot, tot_e = out["resident"]["total"], out["e2e"]["total"]
    return {
        "metric": "large-cohort eval sweep patients/sec (note encoder + pool + FAME forward + EDDI / EO / AUROC metrics)",
        "value": P / (tot * 1e-3), "unit": "patients/s", "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": tot,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"eval sweep: {P} patients x U{{1..16}} chunks = {C} chunks of 512 tokens over {world} GPU(s), patients "
                               "sharded by chunk count; collectives: int64 count all-reduce + logit all-gather (BASELINE configs[4])",
                   "patients": P, "chunks": C, "l2": "inputs (GBs of token ids) stream once; far larger than L2"},
        "stages_ms": out["resident"], "chunks_per_s": C / (out["resident"]["note_encoder_and_pool"] * 1e-3),
        "e2e": {"value": P / (tot_e * 1e-3), "unit": "patients/s", "ms_per_step": tot_e, "stages_ms": out["e2e"],
                "h2d_bytes_per_step": sw.h2d_bytes, "d2h_bytes_per_step": 8 * 914 + 64},
        "gpu_launches": launches // 2, "result": out["result"],
    }


# ====================================================================================================== eager baseline
def torch_eager_b200(dev, steps=3):
    """Stock PyTorch eager on the same B200 (SURVEY 8d 'second informative baseline'): transformers.BertModel on the
    256 x 512 note batch (fp32 and bf16 autocast), and the training step as plain torch ops + torch.autograd + torch
    AdamW on CUDA tensors (the oracle's functional statement of the reference modules, fp32 and bf16 autocast)."""
    import numpy as np
    import torch

    from fairmultimodal_b200 import synth
    out = {}

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    try:
        from transformers import BertConfig, BertModel
        sd = {k[len("BioBert."):]: torch.from_numpy(v) for k, v in
              synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), WSEED).items()}
        bert = BertModel(BertConfig(vocab_size=synth.VOCAB))
        bert.load_state_dict(sd, strict=True)
        bert = bert.to(dev).eval()
        co = synth.make_cohort(CHUNKS // 4, lab_tokens=4, chunks="fixed4", seq_len=SEQ, seed=1234)
        ids, mask = torch.from_numpy(co["input_ids"]).to(dev), torch.from_numpy(co["attention_mask"]).to(dev)
        for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            with torch.no_grad(), ctx:
                ms = timed(lambda: bert(input_ids=ids, attention_mask=mask).last_hidden_state[:, 0, :], steps)
            out["note_encoder_" + name] = {"ms_per_step": ms, "chunks_per_s": CHUNKS / (ms * 1e-3),
                                           "module": "transformers.BertModel (stock, sdpa), 256 chunks per call"}
        del bert, ids, mask
        torch.cuda.empty_cache()
    except Exception as e:
        out["note_encoder_error"] = f"{type(e).__name__}: {e}"
    try:
        from oracle import fame_oracle as O
        shapes = synth.fame_shapes(lab_tokens=TRAIN_L)
        sdp = {k: torch.from_numpy(v).to(dev).requires_grad_(True) for k, v in synth.synth_state_dict(shapes, 4).items()}
        co = synth.make_cohort(TRAIN_B, lab_tokens=TRAIN_L, chunks=0, with_tokens=False, seed=77)
        co["text"] = np.random.default_rng(0).standard_normal((TRAIN_B, 768)).astype(np.float32)
        batch = [torch.from_numpy(co[k]).to(dev) for k in KEYS9]
        pw = torch.from_numpy(synth.pos_weight(co["labels"])).to(dev)
        params = list(sdp.values())
        opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=0.01)

        def train_once():
            opt.zero_grad()
            o = O.fame_forward(sdp, batch, (0.33, 0.33, 0.33))
            total, _, _ = O.fame_loss(o["fused_logits"].float(), batch[8], (batch[2], batch[4], batch[5]), sdp["sig_weights"], pw, 0.8, 0.01)
            total.backward()
            torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 1.0)
            opt.step()

        for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            with ctx:
                ms = timed(train_once, steps)
            out["train_step_" + name] = {"ms_per_step": ms, "patients_per_s": TRAIN_B / (ms * 1e-3),
                                         "module": "plain torch ops (the reference modules' arithmetic) + autograd + "
                                                   "clip_grad_norm_ + torch.optim.AdamW on CUDA, dropout off"}
        del sdp, opt, params
        torch.cuda.empty_cache()
    except Exception as e:
        out["train_step_error"] = f"{type(e).__name__}: {e}"
    return out


# ====================================================================================================== driver
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def finish():
        if world > 1:
            from fairmultimodal_b200 import parallel
            sys.stdout.flush()
            barrier()
            parallel.shutdown()

    pk = peaks()
    sampler = ClockSampler(local) if rank == 0 else _NoSampler()
    if args.config == 3:
        line = bench_config3(args, world, rank, dev, barrier, pk, sampler)
    elif args.config == 5:
        line = bench_config5(args, world, rank, dev, barrier, pk, sampler)
    elif args.config == 2:
        note = bench_note_encoder(args, world, rank, dev, barrier, pk, sampler)
        line = dict(note, n_gpus=world, higher_is_better=True, vs_baseline=None, data="synthetic")
    else:
        line = bench_train(args, world, rank, dev, barrier, pk, sampler)
        line.update(n_gpus=world, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic")
        if not args.skip_note_encoder:
            line["note_encoder"] = bench_note_encoder(args, world, rank, dev, barrier, pk, sampler)
    clocks = sampler.stop()
    if rank == 0:
        line["clocks"] = clocks
        if world == 1 and args.config in (2, 4):
            if args.cpu_chunks > 0:
                v, th, dt, desc, kind = cpu_note_chunks_per_s(args.cpu_chunks)
                cb = {"value": v, "unit": "chunks/s", "cores": th, "kind": kind,
                      "sample": f"{args.cpu_chunks} chunks x 512 tokens ({dt:.1f} s), {desc}, one chunk per call as 10_FAME.py:157-169"}
                (line if args.config == 2 else line.get("note_encoder", {}))["cpu_baseline"] = cb
            if args.config == 4 and args.cpu_train_steps > 0:
                tv, tth, tdt = CpuTrainStep().time(args.cpu_train_steps, warm=1)
                line["cpu_baseline"] = {
                    "value": tv, "unit": "patients/s", "cores": tth, "kind": "port",
                    "sample": f"{args.cpu_train_steps} step(s) of 32 patients, L = 542 ({tdt:.1f} s) after 1 warm-up step: oracle "
                              "forward + loss under torch.autograd, clip_grad_norm_, torch AdamW, fp32, dropout off"}
            if not args.skip_eager:
                line["torch_eager_b200"] = torch_eager_b200(dev)
        print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5],
                    help="4: train step headline + note encoder (default); 2: note encoder; 3: towers fwd+bwd at 1024 patients; "
                         "5: 46 k-patient eval sweep")
    ap.add_argument("--cpu-chunks", type=int, default=120, help="chunks timed for the note encoder's cpu_baseline (~10 s)")
    ap.add_argument("--cpu-train-steps", type=int, default=3, help="CPU training steps timed beside the GPU step (0 = skip)")
    ap.add_argument("--ref-chunks-per-step", type=int, default=16,
                    help="--impl reference --config 2: chunks per step (bounded sample of the 256-chunk step)")
    ap.add_argument("--skip-note-encoder", action="store_true", help="config 4 without the note-encoder sub-object")
    ap.add_argument("--skip-eager", action="store_true", help="skip the stock-PyTorch-eager-on-B200 baselines")
    ap.add_argument("--no-dropout", action="store_true", help="training step with every dropout probability 0")
    ap.add_argument("--c3-patients", type=int, default=1024)
    ap.add_argument("--c5-patients", type=int, default=46000)
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
