"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2 (no GPU): patient sharding with CSR
re-basing, the SUM all-reduce of the loss statistics / evaluation counts, and variable-length all-gather.  The
per-rank compute is stood in for by the CPU oracle (test infrastructure); the product kernels need a B200."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fairmultimodal_b200 import parallel, synth
from oracle import fame_oracle as O


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 32, 1000):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_chunk_balanced_patient_shards_and_rebase():
    co = synth.make_cohort(500, lab_tokens=4, chunks="u1_16", seq_len=16, seed=2)
    offs = co["chunk_offsets"]
    for w in (1, 2, 4, 8):
        sh = parallel.shard_patients_by_chunks(offs, w)
        assert sh[0][0] == 0 and sh[-1][1] == 500 and all(sh[i][1] == sh[i + 1][0] for i in range(w - 1))
        loads = [int(offs[b] - offs[a]) for a, b in sh]
        assert sum(loads) == int(offs[-1])
        assert max(loads) - min(loads) <= 2 * 16            # within two patients' worth of chunks
        for a, b in sh:
            local, (c0, c1) = parallel.rebase_offsets(offs, a, b)
            assert local[0] == 0 and local[-1] == c1 - c0 and local.dtype == np.int32
            np.testing.assert_array_equal(np.diff(local), np.diff(offs[a:b + 1]))
    # pooling per shard == pooling of the whole cohort (no chunk ever crosses a rank)
    cls = np.random.default_rng(0).standard_normal((int(offs[-1]), 768)).astype(np.float32)
    full = O.pool_patient_notes(cls, offs)
    parts = []
    for a, b in parallel.shard_patients_by_chunks(offs, 4):
        local, (c0, c1) = parallel.rebase_offsets(offs, a, b)
        parts.append(O.pool_patient_notes(cls[c0:c1], local))
    np.testing.assert_array_equal(np.concatenate(parts), full)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 64
        co = synth.make_cohort(B, lab_tokens=4, chunks=0, with_tokens=False, seed=9)
        g = torch.Generator().manual_seed(0)
        z = torch.randn(B, 3, generator=g)
        y = torch.from_numpy(co["labels"])
        attrs = [torch.from_numpy(co[k]) for k in ("age_ids", "ethnicity_ids", "insurance_ids")]
        lo, hi = parallel.shard_range(B, rank, world)
        # (a) loss statistics: per-rank counts / sums, SUM all-reduce == statistics of the global batch
        cnt, sums = O.loss_group_stats(z[lo:hi], y[lo:hi], [a[lo:hi] for a in attrs])
        t_cnt, t_sum = torch.from_numpy(cnt), torch.from_numpy(sums)
        parallel.all_reduce_sum_(t_cnt)
        parallel.all_reduce_sum_(t_sum)
        cnt_all, sums_all = O.loss_group_stats(z, y, attrs)
        ok_cnt = bool((t_cnt.numpy() == cnt_all).all())
        ok_sum = bool(np.allclose(t_sum.numpy(), sums_all, rtol=0, atol=1e-12))
        # (b) variable-length all-gather keeps rank order
        sizes = [parallel.shard_range(B, r, world)[1] - parallel.shard_range(B, r, world)[0] for r in range(world)]
        zg = parallel.all_gather_rows(z[lo:hi], sizes)
        ok_gather = bool(torch.equal(zg, z))
        # (c) evaluation counts: summed per-rank confusion cells == global cells
        pred = (torch.sigmoid(z[:, 0]) > 0.5).numpy().astype(int)
        cells = np.array(O.group_rates(y[lo:hi, 0].numpy(), pred[lo:hi], np.ones(hi - lo, bool))[2])
        tc = torch.from_numpy(cells)
        parallel.all_reduce_sum_(tc)
        ok_cells = tuple(tc.tolist()) == O.group_rates(y[:, 0].numpy(), pred, np.ones(B, bool))[2]
        ret[rank] = (ok_cnt, ok_sum, ok_gather, ok_cells)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_statistics_allreduce():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] == (True, True, True, True), (r, ret[r])


def test_flat_layout_and_gradient_buckets():
    """Flat-buffer layout of the training state: the all-reduce buckets are contiguous, follow the order in which
    the backward completes them, tile the buffer together with the never-reduced (exactly-zero gradient) region, and
    that region holds precisely the query / key projections of the demographic BERT."""
    from fairmultimodal_b200 import synth, train
    shapes = synth.fame_shapes(lab_tokens=542)
    names = [(n, int(np.prod(s))) for n, s in shapes.items() if not n.startswith(train.NO_GRAD_PREFIXES)]
    names.sort(key=lambda x: train._layout_key(x[0]))
    offsets, region, total = train.plan_layout(names)
    assert total >= sum(k for _, k in names) and all(o % 8 == 0 for o in offsets.values())
    cuts = train.plan_buckets(region, total)
    lo, hi = region[train._R_NORED]
    nored = [n for n, o in offsets.items() if lo <= o < hi]
    assert len(nored) == 48 and all(".attention.self.query." in n or ".attention.self.key." in n for n in nored)
    assert hi - lo == 12 * 2 * (768 * 768 + 768)
    def bucket_of(name):
        o = offsets[name]
        return [k for k, (a, b) in cuts.items() if a <= o < b]
    assert bucket_of("fusion_mlp.0.weight") == ["tail"] and bucket_of("text_projector.0.bias") == ["tail"]
    assert bucket_of("behrt_lab.pos_embedding") == ["tail"]
    assert bucket_of("sig_weights") == ["tail"] and bucket_of("behrt_demo.age_embedding.weight") == ["tail"]
    assert bucket_of("behrt_demo.bert.encoder.layer.0.intermediate.dense.weight") == ["tail"]
    first, second = train.DEMO_BUCKET_LAYERS
    assert bucket_of(f"behrt_demo.bert.encoder.layer.{first}.output.dense.weight") == [("demo", first)]
    assert bucket_of("behrt_demo.bert.encoder.layer.11.attention.self.value.weight") == [("demo", first)]
    assert bucket_of(f"behrt_demo.bert.encoder.layer.{first - 1}.output.dense.bias") == [("demo", second)]
    assert bucket_of(f"behrt_demo.bert.encoder.layer.{second}.attention.output.LayerNorm.weight") == [("demo", second)]
    # buckets are listed in the order the backward closes them, each starting where the previous one ended
    spans = list(cuts.values())
    assert spans[0][0] == region[train._R_NORED][1] and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_dropout_sites_follow_the_model_configuration():
    """Host logic of the training-step dropout: probabilities are read from the model exactly where the reference keeps
    them (BertConfig, nn.Dropout, nn.MultiheadAttention.dropout), eval mode and set_dropout(0) switch every site off,
    thresholds are round(p * 65536), seeds differ per site and are stable."""
    from fairmultimodal_b200 import modules, train
    model = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(12), "cpu")
    step = torch.zeros(1, dtype=torch.int32)
    model.train()
    ds = train.DropSites(model, step)
    assert ds.any and ds.p_demo_hidden == 0.1 and ds.p_demo_attn == 0.1 and ds.p_fusion == 0.1
    assert ds.lab == [dict(attn=0.1, d1=0.1, act=0.1, d2=0.1)] * 2
    a, b = ds.site("lab.0.d1", 0.1), ds.site("lab.1.d1", 0.1)
    assert a.thresh16 == 6554 and a.seed != b.seed and a.group_shift == 0 and a.step == step.data_ptr()
    assert ds.site("lab.0.d1", 0.1) is a                          # cached: same struct, same seed
    assert ds.site("demo.3.attn", 0.1, 6).group_shift == 6
    assert ds.site("x", 0.0) is None
    assert abs(train.DropSites.inv_keep(a) - 65536 / (65536 - 6554)) < 1e-12 and train.DropSites.inv_keep(None) == 1.0
    with pytest.raises(ValueError):
        ds.site("bad", 1.0)
    model.behrt_lab.transformer_encoder.layers[1].dropout2.p = 0.25
    model.behrt_demo.bert.config.attention_probs_dropout_prob = 0.0
    ds2 = train.DropSites(model, step)
    assert ds2.lab[1]["d2"] == 0.25 and ds2.p_demo_attn == 0.0 and ds2.key() != ds.key()
    model.eval()
    assert not train.DropSites(model, step).any
    model.train()
    modules.set_dropout(model, 0.0)
    assert not train.DropSites(model, step).any
    modules.set_dropout(model, 0.1)
    assert train.DropSites(model, step).key() == ds.key()


def test_flat_train_state_host_logic(monkeypatch):
    """Host side of the flat training state (device kernels stubbed out): module parameters and gradients alias the
    flat buffers at the planned offsets, parameters without gradient in the reference are left out, an in-place update
    from outside (load_state_dict) is detected through the version counters and refreshes the bf16 shadows exactly
    once, and a parameter that stops aliasing the buffer invalidates the state."""
    from fairmultimodal_b200 import modules, ops_train, train
    calls = {"cast": 0, "transpose": 0}
    monkeypatch.setattr(ops_train, "cast_bf16", lambda *a, **k: calls.__setitem__("cast", calls["cast"] + 1))
    monkeypatch.setattr(ops_train, "transpose_bf16_table", lambda *a, **k: calls.__setitem__("transpose", calls["transpose"] + 1))
    model = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(12), "cpu")
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    st = train.get_state(model)
    assert calls == {"cast": 1, "transpose": 1}
    assert train.get_state(model) is st and st.aliased()
    named = dict(model.named_parameters())
    assert not any(n.startswith(train.NO_GRAD_PREFIXES) for n in st.offsets)
    assert set(st.offsets) == {n for n in named if not n.startswith(train.NO_GRAD_PREFIXES)}
    for n in ("sig_weights", "behrt_lab.pos_embedding", "fusion_mlp.3.bias", "behrt_demo.bert.encoder.layer.5.output.dense.weight"):
        o, k = st.offsets[n], named[n].numel()
        assert named[n].data_ptr() == st.p[o:o + k].data_ptr() and named[n].grad.data_ptr() == st.g[o:o + k].data_ptr()
        assert torch.equal(named[n].detach().reshape(-1), st.p[o:o + k]) and torch.equal(st.f(n), sd0[n])
    assert model.classifier_demo.weight.grad is None               # outside the loss path: untouched, as torch leaves it
    st.sync_external_updates()
    assert calls["cast"] == 1                                       # nothing changed
    sd1 = {k: v + 1.0 for k, v in sd0.items()}
    model.load_state_dict(sd1)                                      # in place: the views still alias the flat buffer
    assert st.aliased() and torch.equal(st.f("sig_weights"), sd1["sig_weights"])
    st.sync_external_updates()
    st.sync_external_updates()
    assert calls["cast"] == 2 and calls["transpose"] == 2           # shadows refreshed exactly once
    model.sig_weights.data = model.sig_weights.data.clone()         # stops aliasing -> a new state is built
    assert not st.aliased()
    assert train.get_state(model) is not st
