"""CPU checks of the workload models (fp32 reference paths; the fused CUDA paths are in test_gpu_ops.py)."""
import torch

from adapcc_b200.models.gpt2 import GPT2Config, GPT2DoubleHeads, lm_rows_needed, synthetic_batch


def _grads(m):
    return torch.cat([p.grad.flatten() for p in m.parameters()])


def test_gpt2_double_heads_loss_matches_plain_cross_entropy():
    torch.manual_seed(0)
    cfg = GPT2Config.tiny()
    m = GPT2DoubleHeads(cfg)
    b = synthetic_batch(2, 2, 32, cfg.vocab_size)
    loss, lm, mc = m(**b)
    B, C, T = b["input_ids"].shape
    h = m.hidden(b["input_ids"].reshape(B * C, T), b["token_type_ids"].reshape(B * C, T))
    logits = (h @ m.wte.weight.t())[..., :cfg.vocab_size]
    ref_lm = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, cfg.vocab_size),
                                               b["lm_labels"].reshape(B * C, T)[:, 1:].reshape(-1), ignore_index=-100)
    assert torch.allclose(lm, ref_lm, atol=1e-5)
    assert torch.isfinite(mc) and torch.allclose(loss, lm + mc)


def test_lm_head_on_scored_rows_only_is_equivalent():
    torch.manual_seed(0)
    cfg = GPT2Config.tiny()
    m = GPT2DoubleHeads(cfg)
    b = synthetic_batch(2, 2, 32, cfg.vocab_size)
    need = lm_rows_needed(b["lm_labels"], multiple=8)
    assert need == 16                                    # 2 dialogues x 8 reply tokens of the last candidate
    loss0, _, _ = m(**b)
    loss0.backward()
    g0 = _grads(m).clone()
    for cap in (need, need + 8):                         # exact fit and with padding rows
        m.zero_grad()
        m.lm_row_capacity = cap
        loss1, _, _ = m(**b)
        loss1.backward()
        assert torch.allclose(loss0, loss1, atol=1e-6)
        assert torch.allclose(g0, _grads(m), atol=1e-6)
    m.lm_row_capacity = need - 8                         # too small: loud, not a silent subset
    assert torch.isnan(m(**b)[0])


def test_vit_steps_on_cpu():
    from adapcc_b200.models.vit import ViT, ViTConfig
    torch.manual_seed(0)
    vit = ViT(ViTConfig.tiny())
    x = torch.randn(2, 3, vit.cfg.image_size, vit.cfg.image_size)
    out = vit(x)
    assert out.shape == (2, vit.cfg.num_classes)
    out.sum().backward()
    assert all(p.grad is not None for p in vit.parameters())


def test_zero1_shards_tile_every_bucket_like_the_kernel_partition():
    from adapcc_b200.parallel.engine import shard_of

    for start, end in ((0, 8), (64, 64 + 8 * 1001), (8, 8 + 8 * 7), (0, 8 * 64)):
        for world in (2, 3, 4, 8):
            shards = [shard_of(start, end, r, world, 8) for r in range(world)]
            assert shards[0][0] == start and max(hi for _, hi in shards) == end
            pos = start
            for lo, hi in shards:                      # contiguous, ordered, 16-byte aligned, no overlap
                assert lo == min(pos, end) and hi >= lo and (lo - start) % 8 == 0
                pos = hi if hi > lo else pos
            # same formula as Partition::slice_begin / slice_count, in packs
            packs = (end - start) // 8
            pps = -(-packs // world)
            for r, (lo, hi) in enumerate(shards):
                assert (hi - lo) // 8 == max(0, min(pps, packs - r * pps))


def test_grad_sink_rejects_a_second_gradient_in_one_step():
    import pytest

    from adapcc_b200.parallel.engine import _GradSink

    class _Eng:
        ready = []

        def _grad_ready(self, i):
            self.ready.append(i)

    eng = _Eng()
    sink = _GradSink(eng, 3)
    assert sink.begin() is True
    sink.done()
    assert eng.ready == [3]
    with pytest.raises(RuntimeError, match="second gradient"):
        sink.begin()
    sink.written = False                                # the engine resets sinks at the start of a step
    assert sink.begin() is True


def test_moe_mlp_matches_a_per_token_reference_on_cpu():
    """MoEMLP (top-k softmax gate, capacity-padded batched expert GEMMs, un-sort, weighted combine) against the obvious
    per-token loop; capacity chosen large enough that nothing is dropped, then a tiny capacity to see the drop rule."""
    import torch.nn.functional as F

    from adapcc_b200.models.moe import MoEMLP

    torch.manual_seed(0)
    for top_k in (1, 2):
        moe = MoEMLP(num_expert=4, d_model=16, d_hidden=32, top_k=top_k, capacity_factor=8.0)
        x = torch.randn(37, 16)
        y = moe(x)
        logits = moe.gate(x).float()
        w, idx = torch.topk(F.softmax(logits, -1), top_k, -1)
        if top_k > 1:
            w = w / w.sum(-1, keepdim=True)
        ref = torch.zeros_like(x)
        for t in range(x.shape[0]):
            for j in range(top_k):
                e = int(idx[t, j])
                h = F.gelu(x[t] @ moe.w1[e] + moe.b1[e, 0])
                ref[t] += w[t, j] * (h @ moe.w2[e] + moe.b2[e, 0])
        assert torch.allclose(y, ref, atol=1e-5), (top_k, (y - ref).abs().max())
        y.sum().backward()
        assert all(p.grad is not None for p in moe.parameters())
    # capacity: with factor 0 every expert keeps 8 rows (the minimum); tokens beyond that contribute zero
    moe = MoEMLP(num_expert=2, d_model=16, d_hidden=32, top_k=1, capacity_factor=0.0)
    x = torch.randn(64, 16)
    y = moe(x)
    kept = (y.abs().sum(-1) > 0).sum().item()
    assert kept <= 2 * moe.capacity(64) and kept < 64


def test_flat_engine_state_dict_roundtrip_on_cpu():
    """Checkpoint / resume of the flat engine: per-parameter-name fp32 state, restored in place into an engine with a
    different bucket layout; bf16 parameters are re-derived from the master weights."""
    from adapcc_b200.parallel.engine import FlatDataParallel

    torch.manual_seed(0)
    cfg = GPT2Config.tiny()
    a = FlatDataParallel(GPT2DoubleHeads(cfg), None, world_size=1, lr=3e-4, bucket_mb=0.05)
    a.master.normal_()                                   # pretend some training happened
    a.exp_avg.uniform_(-1, 1)
    a.exp_avg_sq.uniform_(0, 1)
    a.steps_done = 17
    a.step_t.fill_(17)
    sd = a.state_dict()
    assert sd["steps_done"] == 17 and set(sd["params"]) == {n for n, _ in a.model.named_parameters()}
    torch.manual_seed(1)
    b = FlatDataParallel(GPT2DoubleHeads(cfg), None, world_size=1, lr=1e-3, bucket_mb=1.0)
    assert len(b.buckets) != len(a.buckets)
    ptr = b.flat_param.data_ptr()
    b.load_state_dict(sd)
    assert b.flat_param.data_ptr() == ptr and b.steps_done == 17 and int(b.step_t.item()) == 17 and b.lr == 3e-4
    for (na, pa), (nb, pb) in zip(a.model.named_parameters(), b.model.named_parameters()):
        ea, eb = sd["params"][na], b.state_dict()["params"][nb]
        assert torch.equal(ea["master"], eb["master"]) and torch.equal(ea["exp_avg_sq"], eb["exp_avg_sq"])
        assert torch.equal(pb.detach().float(), ea["master"].to(pb.dtype).float())       # params follow the masters
    bad = dict(sd, params={k: v for k, v in list(sd["params"].items())[1:]})
    import pytest
    with pytest.raises(KeyError):
        b.load_state_dict(bad)
    b.load_state_dict(bad, strict=False)


def test_gpt2_sampling_respects_top_k_top_p_and_eos():
    from adapcc_b200.models.gpt2 import sample_sequence

    torch.manual_seed(0)
    cfg = GPT2Config.tiny()
    m = GPT2DoubleHeads(cfg).eval()
    prompt = torch.randint(0, cfg.vocab_size, (1, 9))
    tt = torch.full_like(prompt, cfg.vocab_size - 4)
    g = torch.Generator().manual_seed(1)
    out = sample_sequence(m, prompt, tt, max_new_tokens=7, top_k=5, top_p=0.9, reply_type=cfg.vocab_size - 5, generator=g)
    assert out.shape == (1, 7) and int(out.max()) < cfg.vocab_size
    # top_k = 1 is greedy decoding: deterministic and equal to the arg-max continuation
    a = sample_sequence(m, prompt, tt, max_new_tokens=4, top_k=1, top_p=1.0)
    b = sample_sequence(m, prompt, tt, max_new_tokens=4, top_k=1, top_p=1.0)
    assert torch.equal(a, b)
    h = m.hidden(prompt, tt)
    first = int((h[:, -1].float() @ m.wte.weight.float().t())[:, :cfg.vocab_size].argmax())
    assert int(a[0, 0]) == first
    # stops at EOS
    c = sample_sequence(m, prompt, tt, max_new_tokens=10, top_k=1, top_p=1.0, eos_token=first)
    assert c.shape == (1, 1)
    # longer than the context window: the window slides
    long = sample_sequence(m, torch.randint(0, cfg.vocab_size, (1, cfg.n_positions)), None, max_new_tokens=3, top_k=1, top_p=1.0)
    assert long.shape == (1, 3)
