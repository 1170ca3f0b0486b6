"""GPU parity tests of the decode step's own kernels, through the C ABI, vs plain torch fp32:
self-attention over the paged cache fused with the QKV split-K reduction, and the fused logit rules + top-k."""
import numpy as np
import pytest
import torch

from realtime_whisper_asr_b200 import _lib

pytestmark = pytest.mark.gpu

SF_RULES, SF_FIRST, SF_SUP_ALL_TS, SF_SUP_TEXT = 1, 2, 4, 8


def _self_attention_case(lib, dev, H, G, n_streams, hist, shared, n_splits, idle=(), ragged=False, seed=0):
    """n_streams groups of G rows; every live row sees `hist` cached positions of which the first `shared` are
    common to the rows of its stream (the same slots), then its own new token."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    d = H * 64
    rows = n_streams * G
    if ragged:
        rows -= 2  # the last group is short
    max_len = hist + 3
    n_slots = n_streams * (shared + G * (hist - shared + 1)) + 8
    kpool = (torch.randn(n_slots, d, generator=g) * 0.7).to(torch.bfloat16)
    vpool = (torch.randn(n_slots, d, generator=g) * 0.7).to(torch.bfloat16)
    table = torch.zeros(rows, max_len, dtype=torch.int32)
    kvlen = torch.zeros(rows, dtype=torch.int32)
    slot = torch.zeros(rows, dtype=torch.int32)
    nxt = 0
    hists = []
    for b in range(n_streams):
        common = list(range(nxt, nxt + shared))
        nxt += shared
        for k in range(G):
            r = b * G + k
            if r >= rows:
                continue
            h_r = 0 if r in idle else (hist - (k % 3 if ragged and hist > 4 else 0))
            own = list(range(nxt, nxt + max(h_r - shared, 0)))
            nxt += len(own)
            tab = (common + own)[:h_r]
            table[r, :len(tab)] = torch.tensor(tab, dtype=torch.int32) if tab else table[r, :0]
            kvlen[r] = h_r + 1
            slot[r] = nxt
            nxt += 1
            hists.append(h_r)
    assert nxt <= n_slots
    trow = torch.randperm(rows, generator=g).to(torch.int32)  # page-table rows are not the batch rows
    table_p = torch.zeros_like(table)
    table_p[trow.long()] = table
    partial = torch.randn(n_splits, rows, 3 * d, generator=g) * 0.4
    bias = torch.randn(3 * d, generator=g) * 0.2
    out = torch.zeros(rows, d, dtype=torch.bfloat16, device=dev)
    kp, vp = kpool.to(dev), vpool.to(dev)
    args = [t.to(dev) for t in (partial, bias, slot, table_p, trow, kvlen)]
    _lib.check(lib.b200w_step_self_attention(args[0].data_ptr(), n_splits, args[1].data_ptr(), rows, d, H,
                                             args[2].data_ptr(), kp.data_ptr(), vp.data_ptr(), args[3].data_ptr(),
                                             max_len, args[4].data_ptr(), args[5].data_ptr(), max_len, G,
                                             out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    # reference
    acc = bias.clone().expand(rows, -1).clone()
    for s in range(n_splits):
        acc = acc + partial[s]
    qkv = acc.to(torch.bfloat16)
    q, k_new, v_new = qkv[:, :d].float(), qkv[:, d:2 * d], qkv[:, 2 * d:]
    ref = torch.zeros(rows, d)
    for r in range(rows):
        idx = table[r, :hists[r]].long()
        K = torch.cat([kpool[idx], k_new[r:r + 1]]).float().view(-1, H, 64)
        V = torch.cat([vpool[idx], v_new[r:r + 1]]).float().view(-1, H, 64)
        sc = torch.einsum("hc,thc->ht", q[r].view(H, 64) * 0.125, K)
        ref[r] = torch.einsum("ht,thc->hc", torch.softmax(sc, -1), V).reshape(-1)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 2e-2, f"max abs err {err}"
    # the new keys / values landed in the cache, nothing else moved
    kp_ref, vp_ref = kpool.clone(), vpool.clone()
    kp_ref[slot.long()] = k_new
    vp_ref[slot.long()] = v_new
    assert torch.equal(kp.cpu(), kp_ref) and torch.equal(vp.cpu(), vp_ref)


@pytest.mark.parametrize("H,G,n_streams,hist,shared,n_splits,idle,ragged", [
    (8, 1, 5, 37, 37, 4, (), False),          # greedy: one row per stream
    (8, 5, 4, 100, 80, 4, (), False),         # beams sharing most of their history
    (20, 5, 32, 130, 110, 4, (7, 8, 9), False),  # the bench shape, one stream partly idle
    (4, 5, 3, 64, 0, 2, (), False),           # nothing shared
    (8, 5, 3, 1, 1, 1, (), False),
    (8, 5, 3, 0, 0, 4, (), False),            # first step after an empty prompt: only the own position
    (8, 8, 3, 300, 290, 3, (0,), True),       # ragged histories, short last group
    (12, 3, 7, 33, 20, 5, (4,), True),
])
def test_step_self_attention(lib, cuda_device, H, G, n_streams, hist, shared, n_splits, idle, ragged):
    _self_attention_case(lib, cuda_device, H, G, n_streams, hist, shared, n_splits, idle, ragged, seed=H + G + hist)


def _rules_reference(logits, mask, state, eot, no_ts, tb, max_init, k):
    """torch restatement of the rules (faster-whisper / openai-whisper ApplyTimestampRules + suppress lists)."""
    rows, V = logits.shape
    ids_out = np.full((rows, k), 0x7fffffff, dtype=np.int64)
    lp_out = np.full((rows, k), -np.inf, dtype=np.float32)
    v = torch.arange(V)
    for r in range(rows):
        flags, ts_lo = int(state[r, 0]), int(state[r, 1])
        x = logits[r].clone().double()
        bar = (mask & 1).bool()
        if flags & SF_FIRST:
            bar |= (mask & 2).bool()
        if flags & SF_RULES:
            bar |= v == no_ts
            ts = v >= tb
            if flags & SF_SUP_ALL_TS:
                bar |= ts
            bar |= ts & (v < ts_lo)
            if flags & SF_FIRST:
                bar |= ts & (v > tb + max_init)
                bar |= ~ts
            if flags & SF_SUP_TEXT:
                bar |= (~ts) & (v < eot)
        x[bar] = -np.inf
        lp = torch.log_softmax(x, -1)
        if flags & SF_RULES:
            ts_lp = torch.logsumexp(lp[tb:], -1)
            if ts_lp > lp[:tb].max():
                x[:tb] = -np.inf
                lp = torch.log_softmax(x, -1)
        order = sorted(range(V), key=lambda i: (-lp[i].item(), i))[:k] if V <= 4096 else None
        if order is None:
            top = torch.topk(lp, k + 64)
            cand = sorted(zip((-top.values).tolist(), top.indices.tolist()))[:k]
            order = [i for _, i in cand]
        for j, i in enumerate(order):
            if lp[i] > -np.inf:
                ids_out[r, j] = i
                lp_out[r, j] = lp[i].item()
    return ids_out, lp_out


@pytest.mark.parametrize("V,k,rows", [(51866, 10, 24), (51865, 1, 9), (51864, 16, 5), (2048 + 3, 6, 12)])
def test_sample_topk(lib, cuda_device, V, k, rows):
    g = torch.Generator(device="cpu").manual_seed(V + k)
    ld = (V + 7) // 8 * 8
    tb = V - 1501
    eot, no_ts = tb - 110, tb - 1
    logits = torch.randn(rows, ld, generator=g) * 3.0
    logits[:, tb:V] += 1.5 * torch.randn(rows, 1, generator=g)  # some rows have heavy timestamp mass
    logits[0, 100:140] = logits[0, :V].max() + 0.5  # exact ties inside and at the edge of the top-k
    logits[1, tb + 5] = logits[1, 50] = logits[1].max() + 1.0
    mask = (torch.rand(V, generator=g) < 0.02).to(torch.uint8) | ((torch.rand(V, generator=g) < 0.01).to(torch.uint8) << 1)
    state = torch.zeros(rows, 4, dtype=torch.int32)
    combos = [0, SF_RULES, SF_RULES | SF_FIRST, SF_RULES | SF_SUP_ALL_TS, SF_RULES | SF_SUP_TEXT, SF_FIRST]
    for r in range(rows):
        state[r, 0] = combos[r % len(combos)]
        state[r, 1] = tb + (r * 37) % 1400 if r % 2 else 0
    dl, dm, ds = logits.to(cuda_device), mask.to(cuda_device), state.to(cuda_device)
    res = []
    for streaming in (0, 1):
        ids = torch.zeros(rows, k, dtype=torch.int32, device=cuda_device)
        lps = torch.zeros(rows, k, dtype=torch.float32, device=cuda_device)
        work = dl.clone()  # the streaming kernel masks its input in place
        _lib.check(lib.b200w_sample_topk(work.data_ptr(), rows, ld, V, dm.data_ptr(), ds.data_ptr(), eot, no_ts, tb, 50,
                                         k, ids.data_ptr(), lps.data_ptr(), streaming, 0.0, 0,
                                         torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        res.append((ids.cpu().numpy(), lps.cpu().numpy()))
    # the on-chip kernel and the streaming kernel agree bit for bit
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1].view(np.int32), res[1][1].view(np.int32))
    ref_ids, ref_lp = _rules_reference(logits[:, :V], mask, state.numpy(), eot, no_ts, tb, 50, k)
    assert np.array_equal(res[0][0].astype(np.int64), ref_ids)
    fin = np.isfinite(ref_lp)
    assert np.array_equal(fin, np.isfinite(res[0][1]))
    assert np.abs(res[0][1][fin] - ref_lp[fin]).max() <= 2e-5


def test_sample_temperature(lib, cuda_device):
    """Sampling mode: many draws from one row (different draw indices) follow softmax(masked logits / T); the returned
    log-probability is the token's under that distribution; barred tokens are never drawn; draws are reproducible."""
    g = torch.Generator(device="cpu").manual_seed(11)
    V, T, rows = 2048 + 3, 0.7, 8192
    ld = (V + 7) // 8 * 8
    tb = V - 101
    eot, no_ts = tb - 10, tb - 1
    row = torch.randn(ld, generator=g) * 2.0
    mask = (torch.rand(V, generator=g) < 0.1).to(torch.uint8)
    logits = row.expand(rows, ld).contiguous()
    state = torch.zeros(rows, 4, dtype=torch.int32)
    state[:, 2] = torch.arange(rows, dtype=torch.int32)  # draw index
    state[:, 3] = 5
    dl, dm, ds = logits.to(cuda_device), mask.to(cuda_device), state.to(cuda_device)

    def draw(seed):
        ids = torch.zeros(rows, dtype=torch.int32, device=cuda_device)
        lps = torch.zeros(rows, dtype=torch.float32, device=cuda_device)
        _lib.check(lib.b200w_sample_topk(dl.data_ptr(), rows, ld, V, dm.data_ptr(), ds.data_ptr(), eot, no_ts, tb, 50, 1,
                                         ids.data_ptr(), lps.data_ptr(), 0, T, seed,
                                         torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        return ids.cpu().numpy(), lps.cpu().numpy()

    ids, lps = draw(3)
    ids2, _ = draw(3)
    ids3, _ = draw(4)
    assert np.array_equal(ids, ids2) and not np.array_equal(ids, ids3)
    x = row[:V].double().clone()
    x[mask.bool()] = -np.inf
    ref_lp = torch.log_softmax(x / T, -1).numpy()
    assert np.all(np.isfinite(ref_lp[ids])), "a barred token was drawn"
    assert np.abs(lps - ref_lp[ids]).max() <= 2e-4
    # empirical frequencies of the 12 likeliest tokens vs their probabilities (binomial 5 sigma)
    p = np.exp(ref_lp)
    for tok in np.argsort(-p)[:12]:
        f = np.mean(ids == tok)
        assert abs(f - p[tok]) <= 5 * np.sqrt(p[tok] * (1 - p[tok]) / rows) + 1e-4, (tok, f, p[tok])
    # with the timestamp rules on and the timestamps carrying most of the mass, only timestamps are drawn
    logits2 = logits.clone()
    logits2[:, tb:V] += 8.0
    state2 = state.clone()
    state2[:, 0] = SF_RULES
    ids_t = torch.zeros(rows, dtype=torch.int32, device=cuda_device)
    lps_t = torch.zeros(rows, dtype=torch.float32, device=cuda_device)
    _lib.check(lib.b200w_sample_topk(logits2.to(cuda_device).data_ptr(), rows, ld, V, dm.data_ptr(),
                                     state2.to(cuda_device).data_ptr(), eot, no_ts, tb, 50, 1, ids_t.data_ptr(),
                                     lps_t.data_ptr(), 0, T, 3, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert (ids_t.cpu().numpy() >= tb).all()
