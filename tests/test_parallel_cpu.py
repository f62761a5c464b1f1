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


def _adamw_stub(p, g, m, v, sumsq, max_norm, lr, b1, b2, eps, wd, step, grad_norm_out=None, step_dev=None,
                hyper_dev=None, p_bf16=None):
    """CPU stand-in for fame_clip_adamw (clip_grad_norm_ + torch AdamW arithmetic, elementwise) for the host-logic test."""
    t = int(step_dev.item()) if step_dev is not None else step
    norm = float(sumsq.item()) ** 0.5
    coef = min(1.0, max_norm / (norm + 1e-6))
    gg = g * coef
    m.mul_(b1).add_(gg, alpha=1 - b1)
    v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
    p.mul_(1 - lr * wd)
    p.addcdiv_(m / (1 - b1 ** t), (v / (1 - b2 ** t)).sqrt() + eps, value=-lr)
    if p_bf16 is not None:
        p_bf16.copy_(p)
    if grad_norm_out is not None:
        grad_norm_out.fill_(norm)


def _decay_stub(p, lr, wd, hyper_dev=None, p_bf16=None):
    """CPU stand-in for fame_decay_only (the AdamW step of the zero-gradient query / key region)."""
    p.mul_(1 - lr * wd)
    if p_bf16 is not None:
        p_bf16.copy_(p)


class _NoStream:
    def wait_stream(self, s):
        pass


def _sharded_worker(rank, world, port, ret):
    """Host protocol of the sharded optimizer on gloo (kernels replaced by CPU arithmetic): reduce-scatter of the
    SHARDED buckets + all-reduce of the replicated tail + partial-norm all-reduce + AdamW on this rank's ranges +
    all-gather of the bf16 shadows + sync_masters  ==  all-reduce of everything + AdamW on everything."""
    import contextlib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fairmultimodal_b200 import modules, ops_train, train
        ops_train.cast_bf16 = lambda x, y: y.copy_(x)
        ops_train.transpose_bf16_table = lambda *a, **k: None
        ops_train.grad_sumsq = lambda g, out: out.add_((g.double() ** 2).sum())
        ops_train.clip_adamw = _adamw_stub
        ops_train.decay_only = _decay_stub
        train.torch.cuda.stream = lambda s: contextlib.nullcontext()
        train.torch.cuda.current_stream = lambda *a: _NoStream()
        train.torch.cuda.is_current_stream_capturing = lambda: False
        train.FlatTrainState.post_stream = lambda self: _NoStream()
        torch.manual_seed(0)
        model = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(12), "cpu")
        st = train.get_state(model)
        group = dist.group.WORLD
        plan = st.shard_plan(group)
        assert plan == (rank, world)
        p0 = st.p.clone()
        gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
        grads = [torch.randn(st.n, generator=g) * 0.01 for g in gens]
        qlo, qhi, _ = st.grad_buckets()["qk"]
        for g in grads:
            g[qlo:qhi] = 0                                       # query / key weights: exactly zero gradient on every rank
        ok = []
        for step in (1, 2):
            # ---- reference: everything all-reduced, AdamW on everything (single-process arithmetic)
            if step == 1:
                rp, rm, rv = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
                rpb = torch.zeros(st.n, dtype=torch.bfloat16)
            gsum = sum(grads) * step
            ss = torch.zeros(1, dtype=torch.float64)
            ops_train.grad_sumsq(gsum, ss)
            _adamw_stub(rp, gsum, rm, rv, ss, 1.0, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, p_bf16=rpb)
            # ---- protocol under test
            gathers = train.begin_step(st, group)
            for w in gathers.values():
                w.wait()
            if step > 1:
                # gathered shadows == replicated result of the last step (the query / key shadows are never read by the
                # length-1 forward and are gathered by sync_masters only)
                ok.append(bool(torch.equal(st.pb[qhi:], rpb_prev[qhi:])))
            st.g.copy_(grads[rank] * step)
            red = train._GradReducer(st, group)
            for key in st.grad_buckets():
                red.ready(key)
            red.finish()
            st.clip_and_step(1e-3, 0.01, (0.9, 0.999), 1e-8, max_norm=1.0)
            ok.append(abs(st.grad_norm.item() - float(ss.item()) ** 0.5) <= 1e-6 * float(ss.item()) ** 0.5)
            # before the sync only this rank's ranges of the sharded buckets are current
            for key, (lo, hi, kind) in st.grad_buckets().items():
                a, b = st.my_range(key, plan)
                ok.append(bool(torch.allclose(st.p[a:b], rp[a:b], rtol=0, atol=1e-7)))
                if kind != train.REPLICATED and world > 1:
                    other = (lo, a) if rank > 0 else (b, hi)
                    ok.append(not torch.allclose(st.p[other[0]:other[1]], rp[other[0]:other[1]], rtol=0, atol=1e-7))
            rpb_prev = rpb.clone()
        st.sync_masters(group)
        ok.append(bool(torch.allclose(st.p, rp, rtol=0, atol=1e-7)) and bool(torch.allclose(st.m, rm, rtol=0, atol=1e-9))
                  and bool(torch.allclose(st.v, rv, rtol=0, atol=1e-12)))
        ok.append(st.masters_stale is False and "ZeRO-1" in train.describe_parallel(model, group))
        ret[rank] = tuple(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_sharded_optimizer_protocol():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert all(ret[r]) and len(ret[r]) > 10, (r, [i for i, v in enumerate(ret[r]) if not v], len(ret[r]))


def test_flat_layout_and_gradient_buckets():
    """Flat-buffer layout of the training state: the reduction buckets are contiguous, follow the order in which the
    backward completes them, tile the buffer exactly, end on boundaries that split over 2 / 4 / 8 ranks, and hold what
    their kind promises: 'qk' = exactly the query / key projection weights of the demographic BERT (zero gradient, never
    reduced), SHARDED buckets = only large matrices (consumed through bf16 shadows), REPLICATED tail = every tensor a
    kernel reads in fp32."""
    from fairmultimodal_b200 import synth, train
    shapes = synth.fame_shapes(lab_tokens=542)
    names = [(n, int(np.prod(s))) for n, s in shapes.items() if not n.startswith(train.NO_GRAD_PREFIXES)]
    names.sort(key=lambda x: train._layout_key(x[0]))
    offsets, region, total = train.plan_layout(names)
    assert total >= sum(k for _, k in names) and all(o % 8 == 0 for o in offsets.values())
    cuts = train.plan_buckets(region, total)
    lo, hi, kind = cuts["qk"]
    nored = [n for n, o in offsets.items() if lo <= o < hi]
    assert kind == train.LOCAL and len(nored) == 24
    assert all(n.endswith(("attention.self.query.weight", "attention.self.key.weight")) for n in nored)
    assert hi - lo == 12 * 2 * 768 * 768

    def bucket_of(name):
        o = offsets[name]
        return [k for k, (a, b, _) in cuts.items() if a <= o < b]
    for n in ("fusion_mlp.0.weight", "text_projector.0.bias", "behrt_lab.pos_embedding", "sig_weights",
              "behrt_demo.age_embedding.weight", "behrt_demo.bert.encoder.layer.0.intermediate.dense.bias",
              "behrt_demo.bert.encoder.layer.3.attention.self.query.bias", "behrt_demo.bert.embeddings.word_embeddings.weight",
              "behrt_lab.transformer_encoder.layers.1.norm1.weight", "behrt_lab.transformer_encoder.layers.0.linear1.weight",
              "behrt_lab.transformer_encoder.layers.0.self_attn.in_proj_weight"):
        assert bucket_of(n) == ["tail"], n
    first, second = train.DEMO_BUCKET_LAYERS
    assert (first, second) == (7, 0)
    assert bucket_of(f"behrt_demo.bert.encoder.layer.{first}.output.dense.weight") == [("demo", first)]
    assert bucket_of("behrt_demo.bert.encoder.layer.11.attention.self.value.weight") == [("demo", first)]
    assert bucket_of(f"behrt_demo.bert.encoder.layer.{first - 1}.output.dense.weight") == [("demo", second)]
    assert bucket_of("behrt_demo.bert.encoder.layer.0.intermediate.dense.weight") == [("demo", second)]
    assert bucket_of("behrt_lab.transformer_encoder.layers.1.linear2.weight") == [("lab", 1)]
    # sharded buckets hold large matrices only (their fp32 masters live on one rank: nothing may read them in fp32)
    for key, (a, b, kind) in cuts.items():
        if kind == train.SHARDED:
            inside = [n for n, o in offsets.items() if a <= o < b]
            assert inside and all(n.endswith(train._DEMO_BIG + train._LAB_BIG) for n in inside), key
    assert cuts["tail"][2] == train.REPLICATED
    # buckets are listed in the order the backward closes them, each starting where the previous one ended
    spans = [(a, b) for a, b, _ in cuts.values()]
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all((b - a) % 64 == 0 for a, b in spans)
    # the exposed tail is small: the verdict asked for <= 20 MB beyond the large matrices of lab layer 0
    tail_bytes = (cuts["tail"][1] - cuts["tail"][0]) * 4
    lab0 = 4 * (2304 * 768 + 768 * 768 + 2 * 768 * 2048)
    assert tail_bytes - lab0 <= 20e6, tail_bytes
    # a 6-layer demographic BERT (ablation 08) gets its own two buckets
    shapes6 = {n: s for n, s in shapes.items() if ".encoder.layer." not in n or int(n.split(".encoder.layer.")[1].split(".")[0]) < 6}
    names6 = [(n, int(np.prod(s))) for n, s in shapes6.items() if not n.startswith(train.NO_GRAD_PREFIXES)]
    names6.sort(key=lambda x: train._layout_key(x[0]))
    _, region6, total6 = train.plan_layout(names6)
    cuts6 = train.plan_buckets(region6, total6)
    assert [k for k in cuts6] == ["qk", ("demo", 3), ("demo", 0), ("lab", 1), "tail"]


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
