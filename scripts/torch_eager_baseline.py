"""Informative GPU baseline: the third-party modules the reference itself calls -- transformers.BertModel
for the note encoder, the way BioClinicalBERT_FT uses it (10_FAME.py:133-142) -- run by stock PyTorch eager on the same
B200, fp32 and bf16 autocast, on the bench workload (256 chunks x 512 tokens, random-init BERT-base, vocab 28 996).
Neither the oracle nor this repository's kernels are involved; it answers "what does a B200 give the unmodified
module stack?".  One JSON line per precision.

    python scripts/torch_eager_baseline.py [--chunks 256] [--steps 5] [--device cuda] [--layers 12]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import synth  # noqa: E402

FLOP_PER_TOKEN = 188_743_680


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=256)
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--per-call", type=int, default=0, help="chunks per model call (0 = all at once; 1 = as the reference)")
    a = ap.parse_args()
    from transformers import BertConfig, BertModel
    dev = torch.device(a.device)
    torch.manual_seed(0)
    model = BertModel(BertConfig(vocab_size=synth.VOCAB, num_hidden_layers=a.layers)).to(dev).eval()
    co = synth.make_cohort(max(1, a.chunks // 4), lab_tokens=4, chunks="fixed4", seq_len=a.seq, seed=1234)
    ids = torch.from_numpy(co["input_ids"][:a.chunks]).to(dev)
    mask = torch.from_numpy(co["attention_mask"][:a.chunks]).to(dev)
    per = a.per_call or a.chunks

    def step():
        outs = []
        for s in range(0, a.chunks, per):
            outs.append(model(input_ids=ids[s:s + per], attention_mask=mask[s:s + per]).last_hidden_state[:, 0, :])
        return torch.cat(outs)

    for name, ctx in (("fp32", torch.autocast(dev.type, enabled=False)),
                      ("bf16_autocast", torch.autocast(dev.type, dtype=torch.bfloat16))):
        with torch.no_grad(), ctx:
            for _ in range(2):
                step()
            if dev.type == "cuda":
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.steps
            else:
                t0 = time.perf_counter()
                for _ in range(a.steps):
                    step()
                ms = (time.perf_counter() - t0) * 1e3 / a.steps
        cps = a.chunks / (ms * 1e-3)
        print(json.dumps({"impl": "torch_eager_" + name, "module": "transformers.BertModel (stock, sdpa)", "device": str(dev),
                          "chunks": a.chunks, "seq_len": a.seq, "layers": a.layers, "chunks_per_call": per,
                          "ms_per_step": ms, "chunks_per_s": cps,
                          "tflops_model": cps * a.seq * FLOP_PER_TOKEN * a.layers / 12 / 1e12}), flush=True)


if __name__ == "__main__":
    main()
