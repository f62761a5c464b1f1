"""BASELINE configs[4]: large-cohort evaluation sweep -- P patients x U{1..16} note chunks -> note encoder (256-chunk
batches, CLS rows only) -> chunk->patient pool -> FAME model forward -> threshold calibration (101-point F1 sweep) ->
AUROC / AUPRC / F1 / EDDI / Equalized Odds over age / ethnicity / insurance subgroups.

This is the evaluation half of run_experiment (10_FAME.py:726-731 note embeddings, 866-915 calibrate / evaluate / EDDI)
for a patient-sharded cohort: patients are split over the ranks by chunk count (parallel.shard_patients_by_chunks), a
patient's chunks travel with it, and the only collectives are the int64 count all-reduce and the logit all-gather for
the exact rank metrics (parallel.evaluate_sharded).  bench.py --config 5 and scripts/eval_sweep.py drive it.
"""
from __future__ import annotations

import numpy as np
import torch

from . import metrics, modules, ops, parallel, synth

L_TOKENS, CHUNK_BATCH, PATIENT_BATCH, SEQ = 542, 256, 1024, 512
KEYS = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features")


class EvalSweep:
    def __init__(self, patients, world=1, rank=0, device="cuda", group=None, bert_seed=7, fame_seed=4, cohort_seed=1234):
        self.P, self.world, self.rank, self.dev, self.group = patients, world, rank, torch.device(device), group
        meta = synth.make_cohort(patients, lab_tokens=L_TOKENS, chunks="u1_16", with_tokens=False, seed=cohort_seed)
        self.offs_all = meta["chunk_offsets"]
        self.p_lo, self.p_hi = parallel.shard_patients_by_chunks(self.offs_all, world)[rank]
        self.offs, (c_lo, c_hi) = parallel.rebase_offsets(self.offs_all, self.p_lo, self.p_hi)
        self.C = C = c_hi - c_lo
        # tokens of this rank's chunks only: [CLS] body [SEP], the last chunk of a patient is short
        rng = np.random.default_rng(99 + rank)
        ids = rng.integers(1000, synth.VOCAB, (C, SEQ)).astype(np.int64)
        n_per = np.diff(self.offs)
        valid = np.full(C, SEQ, dtype=np.int64)
        valid[self.offs[1:][n_per > 0] - 1] = rng.integers(16, SEQ + 1, int((n_per > 0).sum()))
        pos = np.arange(SEQ)[None, :]
        ids[:, 0] = synth.CLS_ID
        ids[np.arange(C), valid - 1] = synth.SEP_ID
        ids = np.where(pos < valid[:, None], ids, 0)
        mask = (pos < valid[:, None]).astype(np.int64)
        self.ids_h, self.mask_h = torch.from_numpy(ids).pin_memory(), torch.from_numpy(mask).pin_memory()
        self.ids_d = self.mask_d = None
        sd = {k: torch.from_numpy(v) for k, v in
              synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), bert_seed).items()}
        self.enc = modules.BioClinicalBERT_FT.from_state_dict(sd).to(self.dev)
        del sd
        torch.manual_seed(0)
        fame = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5),
                                                          modules.BEHRTModel_Lab(L_TOKENS), self.dev)
        fame.load_state_dict({k: torch.from_numpy(v) for k, v in
                              synth.synth_state_dict(synth.fame_shapes(lab_tokens=L_TOKENS), fame_seed).items()})
        self.fame = fame.to(self.dev).eval()
        self.shard = {k: torch.from_numpy(meta[k][self.p_lo:self.p_hi]).to(self.dev) for k in KEYS + ("labels",)}
        self.offs_d = torch.from_numpy(self.offs).to(self.dev)
        self.h2d_bytes = int(self.ids_h.numel() * 8 + self.mask_h.numel() * 8)

    @property
    def chunks_total(self):
        return int(self.offs_all[-1])

    def make_resident(self):
        self.ids_d, self.mask_d = self.ids_h.to(self.dev), self.mask_h.to(self.dev)

    def warm(self):
        n = min(self.C, CHUNK_BATCH)
        if n:
            self.enc(self.ids_h[:n].to(self.dev), self.mask_h[:n].to(self.dev))

    def encode_and_pool(self, resident=False):
        C, dev = self.C, self.dev
        cls = torch.empty((C, 768), device=dev, dtype=torch.float32)
        for s in range(0, C, CHUNK_BATCH):
            e = min(C, s + CHUNK_BATCH)
            if resident:
                cls[s:e] = self.enc(self.ids_d[s:e], self.mask_d[s:e])
            else:
                cls[s:e] = self.enc(self.ids_h[s:e].to(dev, non_blocking=True), self.mask_h[s:e].to(dev, non_blocking=True))
        return modules.pool_chunks(cls, self.offs_d)

    def model_forward(self, text):
        outs = []
        n = self.p_hi - self.p_lo
        with torch.no_grad():
            for s in range(0, n, PATIENT_BATCH):
                e = min(n, s + PATIENT_BATCH)
                outs.append(self.fame(*[self.shard[k][s:e] for k in KEYS], text[s:e])["fused_logits"])
        return torch.cat(outs) if outs else torch.empty((0, 3), device=self.dev)

    def metrics_pass(self, logits):
        attrs = [self.shard["age_ids"], self.shard["ethnicity_ids"], self.shard["insurance_ids"]]
        labels = self.shard["labels"].float()
        sweep = torch.from_numpy(metrics._SWEEP).to(self.dev)
        vec = ops.eval_counts(logits, labels, attrs, (0.5, 0.5, 0.5), sweep=sweep)
        parallel.all_reduce_sum_(vec, self.group)
        th = metrics.thresholds_from_hist(metrics.Counts(vec).hist)             # calibrate_thresholds on the cohort
        return th, parallel.evaluate_sharded(logits, labels, attrs, th, group=self.group, verbose=False)

    def run(self, resident=False):
        """One full sweep.  Returns (result dict, {stage: device ms on THIS rank})."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        text = self.encode_and_pool(resident)
        ev[1].record()
        logits = self.model_forward(text)
        ev[2].record()
        th, (m, fair, eddi) = self.metrics_pass(logits)                         # ends with a D2H of counts / rank sums
        ev[3].record()
        torch.cuda.synchronize()
        ms = {"note_encoder_and_pool": ev[0].elapsed_time(ev[1]), "fame_model_forward": ev[1].elapsed_time(ev[2]),
              "thresholds_and_metrics": ev[2].elapsed_time(ev[3]), "total": ev[0].elapsed_time(ev[3])}
        res = {"thresholds": th, "auroc": {k: v["aucroc"] for k, v in m.items()}, "eddi_overall": eddi["overall"],
               "eo_overall": {k: v["overall_eo"] for k, v in fair.items()}}
        return res, ms
