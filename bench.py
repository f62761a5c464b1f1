"""bench.py -- FAME hot-path benchmark (driver contract: one JSON line on stdout from rank 0).

Workload at every N: BASELINE.json configs[1], the BioClinicalBERT (BERT-base, vocab 28 996) note-chunk encoder
forward over 256 chunks x 512 tokens in bf16 per GPU, followed by the chunk->patient mean pool (4 chunks per
patient).  One step = one such batch.  Patients/chunks shard across ranks with no data-path collective
("weak" scaling: per-GPU work fixed).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

  value : chunks/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e   : chunks/s through the public module API from pinned HOST buffers (H2D of ids/mask and D2H of the pooled
          patient embeddings inside the timed region)
  --impl reference : the reference's CPU path for the same step (oracle port of BioClinicalBERT_FT.forward called
          one chunk at a time exactly like 10_FAME.py:157-169, fp32, all host threads), bounded sample per step.
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
WORKLOAD = "note_encoder_fwd_bert_base_256chunks_x_512tok_bf16 + chunk->patient mean pool (BASELINE configs[1])"
METRIC = "512-tok note chunks/sec (BioClinicalBERT note-chunk encoder forward)"
FLOP_PER_TOKEN = 188_743_680           # 12*(2*(4*768^2 + 2*768*3072) + 4*512*768)
WSEED = 7


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Host threads available to this process.  torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the
    reference arm on one core: the CPU legs set the thread count explicitly instead."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_chunks_per_s(n_chunks, threads=None):
    """The reference's CPU path for this workload: BioClinicalBERT_FT.forward on ONE chunk per call (10_FAME.py:
    157-169), fp32, eager -- as restated by oracle/fame_oracle.py (the reference script itself cannot travel to
    the GPU box).  Returns (chunks/s, threads, seconds)."""
    import torch

    from fairmultimodal_b200 import synth
    from oracle import fame_oracle as O

    torch.set_num_threads(threads or host_threads())
    sd = {k: torch.from_numpy(v) for k, v in
          synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), WSEED).items()}
    co = synth.make_cohort(max(1, (n_chunks + 3) // 4), lab_tokens=4, chunks="fixed4", seq_len=SEQ, seed=1234)
    ids, mask = torch.from_numpy(co["input_ids"]), torch.from_numpy(co["attention_mask"])
    with torch.no_grad():
        O.note_cls(sd, ids[:1, :64], mask[:1, :64])            # warm the thread pool
        t0 = time.perf_counter()
        for j in range(n_chunks):
            O.note_cls(sd, ids[j:j + 1], mask[j:j + 1])
        dt = time.perf_counter() - t0
    return n_chunks / dt, torch.get_num_threads(), dt


def cpu_reference_train_patients_per_s(steps=1, threads=None):
    """The reference's CPU path for the training step (10_FAME.py:401-449 on the CPU device): forward of the three
    modules + BCE / LEDDI loss + autograd backward + clip_grad_norm_(1.0) + AdamW, fp32 eager, 32 patients, L = 542 --
    the oracle's forward / loss under torch.autograd, torch's own clip and AdamW.  Dropout off (the oracle restates
    the eval-mode arithmetic), which only makes the CPU side cheaper.  Returns (patients/s, threads, seconds)."""
    import numpy as np
    import torch

    from fairmultimodal_b200 import synth
    from oracle import fame_oracle as O

    torch.set_num_threads(threads or host_threads())
    shapes = synth.fame_shapes(lab_tokens=TRAIN_L)
    sd = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in synth.synth_state_dict(shapes, 4).items()}
    co = synth.make_cohort(TRAIN_B, lab_tokens=TRAIN_L, chunks=0, with_tokens=False, seed=77)
    co["text"] = np.random.default_rng(0).standard_normal((TRAIN_B, 768)).astype(np.float32)
    keys = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids",
            "lab_features", "text", "labels")
    batch = [torch.from_numpy(co[k]) for k in keys]
    pw = torch.from_numpy(synth.pos_weight(co["labels"]))
    params = list(sd.values())
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=0.01)
    t0 = time.perf_counter()
    for _ in range(steps):
        opt.zero_grad()
        o = O.fame_forward(sd, batch, (0.33, 0.33, 0.33))
        total, _, _ = O.fame_loss(o["fused_logits"], batch[8], (batch[2], batch[4], batch[5]), sd["sig_weights"], pw,
                                  0.8, 0.01)
        total.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 1.0)
        opt.step()
    dt = time.perf_counter() - t0
    return steps * TRAIN_B / dt, torch.get_num_threads(), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = args.ref_chunks_per_step                       # bounded sample of the 256-chunk step
    if args.warmup > 0:
        cpu_reference_chunks_per_s(min(args.warmup, 3))        # untimed warm-up chunks (thread pool, allocator)
    v, threads, dt = cpu_reference_chunks_per_s(per_step * args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "chunks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{per_step} chunk(s) per step, one chunk per call"},
        "cpu_baseline": {"value": v, "unit": "chunks/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step * args.steps} chunks x 512 tokens, fp32, oracle port of "
                                   "BioClinicalBERT_FT.forward, one chunk per call"},
        "e2e": {"value": v, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


TRAIN_B, TRAIN_L = 32, 542


def bench_train(args, world, rank, dev, barrier):
    """FAME training step (forward + BCE/LEDDI loss + backward + clip + AdamW), 32 patients per GPU, L = 542."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from fairmultimodal_b200 import modules, synth, train

    torch.manual_seed(0)
    demo = modules.BEHRTModel_Demo(5, 2, 5, 5)
    lab = modules.BEHRTModel_Lab(TRAIN_L)
    model = modules.MultimodalTransformer_EDDI_Sigmoid(768, demo, lab, dev).to(dev)
    if args.no_dropout:
        modules.set_dropout(model, 0.0)
    n_batches = 4
    co = synth.make_cohort(TRAIN_B * n_batches, lab_tokens=TRAIN_L, chunks=0, with_tokens=False, seed=77 + rank)
    co["text"] = np.random.default_rng(rank).standard_normal((TRAIN_B * n_batches, 768)).astype(np.float32)
    keys = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids",
            "lab_features", "text", "labels")
    host = [[torch.from_numpy(co[k][i * TRAIN_B:(i + 1) * TRAIN_B]).pin_memory() for k in keys] for i in range(n_batches)]
    devb = [[x.to(dev) for x in b] for b in host]
    pw = torch.from_numpy(synth.pos_weight(co["labels"])).to(dev)
    hp = dict(lr=1e-5, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)
    w = (0.33, 0.33, 0.33)
    group = dist.group.WORLD if world > 1 else None
    model.train()
    out_h = torch.empty(4, dtype=torch.float32).pin_memory()

    def step(i, from_host):
        b = [x.to(dev, non_blocking=True) for x in host[i % n_batches]] if from_host else devb[i % n_batches]
        loss = train.optimisation_step(model, b, pw, 0.8, 0.01, w, hp, group=group)
        if from_host:
            out_h.copy_(loss, non_blocking=True)

    res = {}
    for name, from_host in (("resident", False), ("e2e", True)):
        for i in range(4):
            step(i, from_host)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.train_steps):
            step(i, from_host)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = t.item() / args.train_steps
    st = train.get_state(model)
    n_params = int(st.n)
    train.release_graphs(model)            # before the process group goes away (captured NCCL kernels)
    return {
        "metric": "FAME train patients/sec (BASELINE configs[3]: full training step, 32 patients/GPU, L=542 lab tokens, "
                  "3 tasks, text embeddings precomputed as in 10_FAME.py:729-731)",
        "value": world * TRAIN_B / (res["resident"] * 1e-3), "unit": "patients/s", "ms_per_step": res["resident"],
        "e2e": {"value": world * TRAIN_B / (res["e2e"] * 1e-3), "unit": "patients/s", "ms_per_step": res["e2e"],
                "h2d_bytes_per_step": int(sum(x.numel() * x.element_size() for x in host[0])), "d2h_bytes_per_step": 16},
        "global_batch": world * TRAIN_B, "steps": args.train_steps, "params": n_params,
        "cuda_graph": bool(train.USE_CUDA_GRAPH and (group is None or train._GRAPH_WITH_COLLECTIVES)),
        "dropout": "off (parity configuration)" if args.no_dropout else
                   "0.1 at every site of the reference's train() mode (masks from an in-kernel counter hash)", "dtype": "bf16 GEMMs, fp32 master weights / optimizer",
        "collectives": "none" if world == 1 else "all-reduce(SUM) of 104 int64 loss statistics + flat fp32 gradient buffer",
    }


def run_ours(args):
    import torch
    import torch.distributed as dist

    from fairmultimodal_b200 import modules, ops, synth

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

    if args.only_train:
        info = bench_train(args, world, rank, dev, barrier)
        if rank == 0:
            print(json.dumps({"train": info}), flush=True)
        if world > 1:
            from fairmultimodal_b200 import parallel
            sys.stdout.flush()
            barrier()
            parallel.shutdown()
        return

    # random-init BERT-base (vocab 28 996), deterministic in the seed; every rank holds a replica
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
        h = model.encode_chunks(ids_d[i % n_batches], mask_d[i % n_batches])
        return modules.pool_chunks(h, offs, ldx=SEQ * 768, cols=768)

    def step_e2e(i):
        ids = ids_h[i % n_batches].to(dev, non_blocking=True)
        mask = mask_h[i % n_batches].to(dev, non_blocking=True)
        h = model.encode_chunks(ids, mask)
        out_h.copy_(modules.pool_chunks(h, offs, ldx=SEQ * 768, cols=768), non_blocking=True)

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    barrier()

    # ---- timed region 1: inputs resident in HBM; every launch bracketed by events for the roofline ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCHES
    ops.start_trace()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    barrier()
    trace = ops.stop_trace()
    launches = ops.LAUNCHES - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end from pinned host memory ----
    for i in range(2):
        step_e2e(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()

    # ---- second metric of BASELINE.json: FAME train patients/sec (config 4a: 32 patients per GPU, L = 542 lab
    # tokens, text embeddings precomputed as in the reference; data parallel with the statistic + gradient all-reduce)
    train_info = None
    if not args.skip_train:
        del model, ids_d, mask_d
        torch.cuda.empty_cache()
        train_info = bench_train(args, world, rank, dev, barrier)

    if rank == 0:
        pk = peaks()
        by = {}
        for name, tag, a, b, work in trace:
            d = by.setdefault(name, [0.0, 0.0, 0])
            d[0] += a.elapsed_time(b); d[1] += work; d[2] += 1
        g = by["fame_gemm_bias_act"]
        gemm_tf = g[1] / (g[0] * 1e-3) / 1e12
        kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[2] / args.steps,
                       "share": v[0] / sum(x[0] for x in by.values())} for k, v in by.items()}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("fame_gemm_bias_act")
        value = world * CHUNKS * args.steps / (ms * 1e-3)
        cpu_v, cpu_threads, cpu_dt = cpu_reference_chunks_per_s(args.cpu_chunks) if (world == 1 and args.cpu_chunks > 0) else (None, None, None)
        line = {
            "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chunks_per_gpu_per_step": CHUNKS, "seq_len": SEQ,
                       "patients_per_gpu_per_step": patients, "parallelism": f"dp{world} (chunks sharded, no collective)",
                       "weights": "random-init BERT-base, vocab 28996 (no checkpoint reachable)",
                       "l2": "per-step working set (216 MB bf16 weights + >1 GB activations) exceeds the 126 MB L2; "
                             "input batches rotate between steps"},
            "tflops_model": value * SEQ * FLOP_PER_TOKEN / 1e12,
            "tensor_frac_of_sustained_peak": value / world * SEQ * FLOP_PER_TOKEN / 1e12 / pk["tf_sust"],
            "e2e": {"value": world * CHUNKS * args.steps / (ms_e2e * 1e-3), "unit": "chunks/s",
                    "h2d_bytes_per_step": int(ids_h[0].numel() * 8 + mask_h[0].numel() * 8),
                    "d2h_bytes_per_step": int(out_h.numel() * 4), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"kernel": "gemm_bf16_tcgen05_kernel", "bound": "tensor", "achieved": gemm_tf,
                         "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": gemm_tf / pk["tf_sust"],
                         "traffic": traffic, "peak_source": f"{pk['src']} (sustained cuBLAS bf16)",
                         "launches_per_step": g[2] / args.steps, "avg_launch_ms": g[0] / g[2]},
            "kernels": kernels,
            "clocks": clocks,
        }
        if train_info is not None:
            if world == 1 and args.cpu_train_steps > 0:
                tv, tthreads, tdt = cpu_reference_train_patients_per_s(args.cpu_train_steps)
                train_info["cpu_baseline"] = {
                    "value": tv, "unit": "patients/s", "cores": tthreads, "kind": "port",
                    "sample": f"{args.cpu_train_steps} step(s) of 32 patients, L = 542 ({tdt:.1f} s): oracle forward + "
                              "loss under torch.autograd, clip_grad_norm_, torch AdamW, fp32, dropout off"}
            line["train"] = train_info
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": "chunks/s", "cores": cpu_threads, "kind": "port",
                                    "sample": f"{args.cpu_chunks} chunks x 512 tokens ({cpu_dt:.1f} s), fp32 oracle port "
                                              "of BioClinicalBERT_FT.forward, one chunk per call as 10_FAME.py:157-169"}
        print(json.dumps(line), flush=True)
    if world > 1:
        from fairmultimodal_b200 import parallel
        sys.stdout.flush()
        barrier()
        parallel.shutdown()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-chunks", type=int, default=160, help="chunks timed for the cpu_baseline leg (~10-20 s)")
    ap.add_argument("--ref-chunks-per-step", type=int, default=16,
                    help="--impl reference: chunks per step (bounded sample of the 256-chunk step)")
    ap.add_argument("--skip-train", action="store_true", help="only the note-encoder workload")
    ap.add_argument("--cpu-train-steps", type=int, default=1, help="CPU training steps timed beside the GPU step (0 = skip)")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--no-dropout", action="store_true", help="training step with every dropout probability 0")
    ap.add_argument("--only-train", action="store_true", help="diagnostic: print only the training-step object")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
