"""Generate tests/golden/*.npz by running the UNMODIFIED reference (10_FAME.py) on seeded synthetic inputs.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
Weights are not stored: they are regenerated from fairmultimodal_b200.synth.synth_state_dict(seed) and loaded
into the reference modules with load_state_dict(strict=True), so the fixtures hold inputs and outputs only.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fairmultimodal_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WSEED = 7


def to_torch_sd(sd, strip=""):
    return {k[len(strip):] if strip and k.startswith(strip) else k: torch.from_numpy(v.copy()) for k, v in sd.items()}


class _LogitStub(torch.nn.Module):
    """Stands in for the model inside the reference's eval functions: returns logits smuggled through the
    `aggregated_text_embedding` slot so that calibrate_thresholds / evaluate_model_multi /
    update_dynamic_weights_all_tasks run unmodified on controlled logits."""

    beta = 1.0

    def forward(self, a, b, c, d, e, f, g, text, **kw):
        return {"fused_logits": text[:, 0:3],
                "modality_logits": {"demo": text[:, 3:6], "lab": text[:, 6:9], "text": text[:, 9:12]}}


def make_loader(ref, cohort, text, bsz):
    t = lambda k: torch.from_numpy(cohort[k])
    ds = ref.TensorDataset(t("demo_dummy_ids"), t("demo_attn_mask"), t("age_ids"), t("gender_ids"),
                           t("ethnicity_ids"), t("insurance_ids"), t("lab_features"), torch.from_numpy(text),
                           t("labels"))
    return ref.DataLoader(ds, batch_size=bsz, shuffle=False)


def golden_metrics(ref):
    rng = np.random.default_rng(11)
    N = 1500
    co = synth.make_cohort(N, lab_tokens=4, chunks=0, with_tokens=False, seed=5)
    y = co["labels"]
    # informative but noisy logits, rounded so that float32 sigmoid produces exact ties and threshold hits
    z = (y * 2 - 1) * 0.8 + rng.standard_normal((N, 3)) * 1.5
    z = np.round(z * 4) / 4
    mod = np.round((rng.standard_normal((N, 9)) + np.tile(y, 3) * 0.7) * 4) / 4
    text = np.concatenate([z, mod], axis=1).astype(np.float32)
    loader = make_loader(ref, co, text, 64)
    stub = _LogitStub()
    out = {"logits": text[:, :3], "mod_logits": text[:, 3:], "labels": y, "age": co["age_ids"],
           "eth": co["ethnicity_ids"], "ins": co["insurance_ids"]}
    th = ref.calibrate_thresholds(stub, loader, "cpu")
    out["thresholds"] = np.array([th[k] for k in ("mortality", "los", "mechanical_ventilation")])
    metrics, lg, lb, ag, et, ins, fair = ref.evaluate_model_multi(stub, loader, "cpu", th)
    names = ("mortality", "los", "mechanical_ventilation")
    out["aucroc"] = np.array([metrics[n]["aucroc"] for n in names])
    out["auprc"] = np.array([metrics[n]["auprc"] for n in names])
    out["f1"] = np.array([metrics[n]["f1"] for n in names])
    out["tpr"] = np.array([metrics[n]["TPR"] for n in names], dtype=np.float64)
    out["fpr"] = np.array([metrics[n]["fpr"] for n in names], dtype=np.float64)
    out["precision"] = np.array([metrics[n]["precision"] for n in names])
    out["eo"] = np.array([[fair[n][a]["eo_metric"] for a in ("age", "ethnicity", "insurance")] for n in names])
    out["tpr_diff"] = np.array([[fair[n][a]["avg_tpr_diff"] for a in ("age", "ethnicity", "insurance")] for n in names])
    out["fpr_diff"] = np.array([[fair[n][a]["avg_fpr_diff"] for a in ("age", "ethnicity", "insurance")] for n in names])
    out["overall_eo"] = np.array([fair[n]["overall_eo"] for n in names])
    # EDDI tail of run_experiment (10_FAME.py:887-915)
    ed = []
    for i, n in enumerate(names):
        probs = torch.sigmoid(torch.tensor(lg))[:, i].numpy().squeeze()
        row = []
        for a, gl in ((ag, [0, 1, 2, 3]), (et, [0, 1, 2, 3, 4]), (ins, [0, 1, 2, 3, 4, 5])):
            row.append(ref.compute_eddi(lb[:, i], probs, np.array(a), threshold=th[n], complete_groups=gl)[0])
        ed.append(row)
    out["eddi"] = np.array(ed)
    # compute_eddi without complete_groups and at fixed 0.5
    e05 = []
    for i in range(3):
        probs = torch.sigmoid(torch.tensor(lg))[:, i].numpy()
        e05.append([ref.compute_eddi(lb[:, i], probs, np.array(a))[0] for a in (ag, et, ins)])
    out["eddi_unique_groups_t05"] = np.array(e05)
    # weight update (10_FAME.py:315-399) for two consecutive epochs
    w0 = {n: {"demo": 0.33, "lab": 0.33, "text": 0.33} for n in names}
    w1 = ref.update_dynamic_weights_all_tasks(stub, loader, "cpu", w0, beta=1.0)
    w2 = ref.update_dynamic_weights_all_tasks(stub, loader, "cpu", w1, beta=1.0)
    out["weights_epoch1"] = np.array([[w1[n][m] for m in ("demo", "lab", "text")] for n in names])
    out["weights_epoch2"] = np.array([[w2[n][m] for m in ("demo", "lab", "text")] for n in names])
    # hand-evaluated known answer from SURVEY.md section 4
    ka = ref.compute_eddi(np.array([0, 1, 1, 0, 1, 0]), np.array([.1, .9, .2, .8, .7, .3]),
                          np.array([0, 0, 1, 1, 2, 2]), complete_groups=[0, 1, 2, 3])
    out["known_answer_eddi"] = np.array([ka[0], ka[1][0], ka[1][1], ka[1][2]])
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)
    print("metrics.npz", {k: v.shape for k, v in out.items()})


def build_ref_model(ref, L):
    demo = ref.BEHRTModel_Demo(5, 2, 5, 5, hidden_size=768)
    lab = ref.BEHRTModel_Lab(lab_token_count=L, hidden_size=768, nhead=8, num_layers=2)
    model = ref.MultimodalTransformer_EDDI_Sigmoid(768, demo, lab, "cpu", fusion_hidden=512, beta=1.0)
    shapes = synth.fame_shapes(lab_tokens=L)
    sd = model.state_dict()
    assert list(sd.keys()) == list(shapes.keys()), "state_dict key order differs from synth.fame_shapes"
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    model.load_state_dict(to_torch_sd(synth.synth_state_dict(shapes, WSEED)), strict=True)
    for m in model.modules():                       # parity runs: dropout off
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    return model


def golden_model(ref):
    L, B = 24, 12
    co = synth.make_cohort(B, lab_tokens=L, chunks=0, with_tokens=False, seed=21)
    rng = np.random.default_rng(3)
    text = (rng.standard_normal((B, 768)) * 0.5).astype(np.float32)
    model = build_ref_model(ref, L)
    model.eval()
    t = lambda k: torch.from_numpy(co[k])
    args = (t("demo_dummy_ids"), t("demo_attn_mask"), t("age_ids"), t("gender_ids"), t("ethnicity_ids"),
            t("insurance_ids"), t("lab_features"), torch.from_numpy(text))
    w = {"mortality": {"demo": 0.41, "lab": 0.27, "text": 0.32}}
    with torch.no_grad():
        o = model(*args, old_eddi_weights=w, return_modality_logits=True, return_gated_vector=True,
                  return_intermediate=True)
        o_default = model(*args)
        demo_emb = model.behrt_demo(*args[:6])
        lab_emb = model.behrt_lab(args[6])
    out = {k: co[k] for k in ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids",
                              "insurance_ids", "lab_features", "labels")}
    out.update(text=text, weights=np.array([0.41, 0.27, 0.32]), wseed=np.array(WSEED),
               fused_logits=o["fused_logits"].numpy(), fused_logits_default=o_default["fused_logits"].numpy(),
               gated_vector=o["gated_vector"].numpy(), fusion_pre_relu=o["fusion_pre_relu"].numpy(),
               sigmoid_weights=o["sigmoid_weights"].numpy(), demo_embedding=demo_emb.numpy(),
               lab_embedding=lab_emb.numpy(),
               **{f"modality_logits_{m}": o["modality_logits"][m].numpy() for m in ("demo", "lab", "text")})
    # one train_step iteration (dropout off) -> losses, gradient norms and a few updated parameters
    model.train()
    pw = torch.from_numpy(synth.pos_weight(co["labels"]))
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=pw)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=0.01)
    loader = make_loader(ref, co, text, B)
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    # record the value torch's own clip_grad_norm_ returns (train_step discards it): wrap the torch function,
    # the reference stays unmodified
    seen = {}
    real_clip = torch.nn.utils.clip_grad_norm_

    def recording_clip(*a, **kw):
        seen["norm"] = real_clip(*a, **kw)
        return seen["norm"]

    torch.nn.utils.clip_grad_norm_ = recording_clip
    tot, bce = ref.train_step(model, loader, opt, "cpu", crit, beta=1.0, lambda_edd=0.8, lambda_l1=0.01,
                              old_eddi_weights=w)
    torch.nn.utils.clip_grad_norm_ = real_clip
    out["train_total_loss"], out["train_bce_loss"], out["pos_weight"] = np.array(tot), np.array(bce), pw.numpy()
    out["preclip_total_grad_norm"] = np.array(float(seen["norm"]))
    gn = {k: (p.grad.norm().item() if p.grad is not None else -1.0) for k, p in model.named_parameters()}
    keys = ["sig_weights", "fusion_mlp.3.weight", "fusion_mlp.0.bias", "lab_projector.0.weight",
            "behrt_lab.pos_embedding", "behrt_lab.transformer_encoder.layers.1.linear2.weight",
            "behrt_lab.transformer_encoder.layers.0.self_attn.in_proj_weight",
            "behrt_demo.bert.encoder.layer.11.output.dense.weight",
            "behrt_demo.bert.encoder.layer.0.attention.self.value.weight",
            "behrt_demo.bert.encoder.layer.0.attention.self.query.weight",
            "behrt_demo.age_embedding.weight", "classifier_demo.weight", "behrt_demo.bert.pooler.dense.weight"]
    out["grad_norm_keys"] = np.array(keys)
    out["grad_norms"] = np.array([gn[k] for k in keys])      # AFTER in-place clipping (10_FAME.py:446)
    out["total_grad_norm"] = np.array(np.sqrt(sum(v * v for v in gn.values() if v >= 0)))
    after = model.state_dict()
    for k in ("sig_weights", "fusion_mlp.3.weight", "behrt_demo.age_embedding.weight"):
        out["delta__" + k] = (after[k] - before[k]).numpy()
    np.savez_compressed(os.path.join(OUT, "model_step.npz"), **out)
    print("model_step.npz written; losses", tot, bce)


def golden_notes(ref):
    from transformers import BertConfig, BertModel

    co = synth.make_cohort(3, lab_tokens=4, chunks="u0_4", seed=9, seq_len=512)
    # shrink the work: keep at most 3 chunks in total, make one patient note-less
    offs = np.array([0, 2, 2, 3], dtype=np.int32)
    ids, mask = co["input_ids"][:3].copy(), co["attention_mask"][:3].copy()
    mask[1, 100:] = 0
    ids[1, 100:] = 0
    ids[1, 99] = synth.SEP_ID
    bert = BertModel(BertConfig(vocab_size=synth.VOCAB))
    shapes = synth.bert_shapes("", synth.VOCAB)
    sd = bert.state_dict()
    assert set(sd.keys()) == set(shapes.keys()), set(sd.keys()) ^ set(shapes.keys())
    bert.load_state_dict(to_torch_sd(synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), WSEED),
                                     strip="BioBert."), strict=True)
    model = ref.BioClinicalBERT_FT(bert, bert.config, "cpu").eval()

    # drive the unmodified apply_bioclinicalbert_on_patient_notes with a tokenizer stub that replays our ids
    import pandas as pd

    class Tok:
        def __init__(self):
            self.i = 0

        def encode_plus(self, text, **kw):
            j = int(text.split("#")[1])
            return {"input_ids": torch.from_numpy(ids[j:j + 1]), "attention_mask": torch.from_numpy(mask[j:j + 1])}

    df = pd.DataFrame({"subject_id": [10, 20, 30], "note_chunk_0": ["chunk#0", None, "chunk#2"],
                       "note_chunk_1": ["chunk#1", "   ", None]})
    pooled = ref.apply_bioclinicalbert_on_patient_notes(df, ["note_chunk_0", "note_chunk_1"], Tok(), model, "cpu")
    with torch.no_grad():
        cls = torch.cat([model(torch.from_numpy(ids[j:j + 1]), torch.from_numpy(mask[j:j + 1])) for j in range(3)])
    np.savez_compressed(os.path.join(OUT, "notes.npz"), input_ids=ids, attention_mask=mask, offsets=offs,
                        cls=cls.numpy(), pooled=pooled.astype(np.float32), wseed=np.array(WSEED))
    print("notes.npz written", cls.shape, pooled.shape)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    ref = ref_loader.load()
    which = sys.argv[1:] or ["metrics", "model", "notes"]
    if "metrics" in which:
        golden_metrics(ref)
    if "model" in which:
        golden_model(ref)
    if "notes" in which:
        golden_notes(ref)
