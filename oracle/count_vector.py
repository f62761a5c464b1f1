"""TEST INFRASTRUCTURE ONLY: numpy restatement of the integer vector written by fame_eval_counts
(fairmultimodal_b200/csrc/metrics.cuh; layout in include/fame_b200.h):

    [0, 288)    confusion cells [outcome][attr][code 0..7][TP, FN, FP, TN]   prediction = (double) p_f32 > thr[outcome]
    [288, 300)  totals          [outcome][TP, FN, FP, TN]
    [300, 912)  F1-sweep hist   [outcome][label][k] = samples with exactly k sweep thresholds strictly below p
    912         number of samples;   913  error flag (a code outside 0..7)

p_f32 = torch.sigmoid(float32 logits) as the reference computes it (10_FAME.py:471-476, 516-518).  Lets the CPU suite
drive the host formulas of fairmultimodal_b200/metrics.py (thresholds, EDDI, EO, weight update) without a GPU; the GPU
suite checks the kernel's vector against per-cell counts of the oracle."""
import numpy as np
import torch


def eval_count_vector(logits, labels, attrs, thr, sweep=None):
    p = torch.sigmoid(torch.as_tensor(np.asarray(logits, dtype=np.float32))).numpy().astype(np.float64)
    y = np.asarray(labels) > 0.5
    v = np.zeros(914, dtype=np.int64)
    conf = v[:288].reshape(3, 3, 8, 4)
    tot = v[288:300].reshape(3, 4)
    hist = v[300:912].reshape(3, 2, 102)
    for o in range(3):
        pred = p[:, o] > float(thr[o])
        cells = (pred & y[:, o], ~pred & y[:, o], pred & ~y[:, o], ~pred & ~y[:, o])      # TP, FN, FP, TN
        for c, m in enumerate(cells):
            tot[o, c] = int(m.sum())
            for a in range(3):
                code = np.asarray(attrs[a])
                for gcode in range(8):
                    conf[o, a, gcode, c] = int((m & (code == gcode)).sum())
        if sweep is not None:
            k = (p[:, o][:, None] > np.asarray(sweep, dtype=np.float64)[None, :]).sum(axis=1)
            for lab in (0, 1):
                hist[o, lab] = np.bincount(k[y[:, o] == bool(lab)], minlength=102)
    v[912] = len(p)
    v[913] = int(any(((np.asarray(a) < 0) | (np.asarray(a) > 7)).any() for a in attrs))
    return v
