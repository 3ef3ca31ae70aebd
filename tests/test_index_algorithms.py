"""CPU models of the three index algorithms behind the deterministic gradient path, restated line by line from
satrans_b200/csrc/rows.cu and bwd_common.cuh and checked on random inputs (the CUDA kernels themselves are checked on the GPU
by the bit-identical-runs and gradient-parity tests):
  * the chunked segmented reduce over the sorted (slot, token) order with boundary pieces and the fix-up pass,
  * the CTA -> tile partition with every CTA inside one scenario,
  * one pass of the stable LSD radix sort (tile histograms, digit-major scan, warp-match ranking)."""
import numpy as np
import pytest

RCH = 32
CI_HEAD, CI_TAIL, CI_SINGLE = 1, 2, 4

def emulate(keys, vals, dx):
    N = len(keys); n_chunks = (N + RCH - 1)//RCH
    out = {}
    pieces = np.zeros((n_chunks, 2, dx.shape[1])); info = np.zeros((n_chunks, 2), dtype=np.int64)
    def rows_out(k, acc):
        assert k not in out, 'double emit %d' % k
        out[k] = acc.copy()
    for g in range(n_chunks):
        i0 = g*RCH; i1 = min(N, i0+RCH)
        cur = keys[i0]; head_cont = i0 > 0 and keys[i0-1] == cur; first_run = True
        acc = np.zeros(dx.shape[1])
        for i in range(i0, i1):
            kk = keys[i]
            if kk != cur:
                if first_run and head_cont: pieces[g,0] = acc
                else: rows_out(cur, acc)
                acc = np.zeros(dx.shape[1]); cur = kk; first_run = False
            acc = acc + dx[vals[i]]
        tail_cont = i1 < N and keys[i1] == cur
        if first_run and head_cont: pieces[g,0] = acc
        elif tail_cont: pieces[g,1] = acc
        else: rows_out(cur, acc)
        info[g] = ((CI_HEAD if head_cont else 0) | (CI_TAIL if tail_cont else 0) | (CI_SINGLE if first_run else 0), cur)
    for g in range(n_chunks):
        fl, key = info[g]
        owns = (fl & CI_TAIL) and not ((fl & CI_SINGLE) and (fl & CI_HEAD))
        if not owns: continue
        acc = pieces[g,1].copy(); done = False; j = g+1
        while not done:
            for u in range(8):
                jj = min(j+u, n_chunks-1)
                if not done:
                    acc = acc + pieces[jj,0]
                    goes_on = (info[jj,0] & CI_SINGLE) and (info[jj,0] & CI_TAIL) and (j+u < n_chunks-1)
                    if not goes_on: done = True
            j += 8
        rows_out(key, acc)
    return out



def test_segmented_reduce_over_sorted_order():
    rng = np.random.default_rng(0)
    for trial in range(60):
        N = int(rng.integers(1, 600))
        nkeys = int(rng.integers(1, 40))
        p = rng.random(nkeys) ** 4
        p /= p.sum()
        _, keys = np.unique(np.sort(rng.choice(nkeys, size=N, p=p)), return_inverse=True)      # every slot present, hot slots
        vals = rng.permutation(N)
        dx = rng.integers(-5, 6, size=(N, 4)).astype(np.float64)
        out = emulate(keys, vals, dx)
        ref = np.zeros((keys.max() + 1, 4))
        np.add.at(ref, keys, dx[vals])
        assert set(out) == set(range(keys.max() + 1))
        assert all(np.array_equal(v, ref[k]) for k, v in out.items())


def scenario_partition(hist, spt, G, c_sel=-1, g_sel=-1):
    T = 0; K = 0
    for h in hist:
        nt = (h + spt - 1)//spt; T += nt; K += nt > 0
    extra = G - K
    cbase = tbase = 0; t0 = t1 = 0; g_of = -1; c0 = c1 = 0
    for g, h in enumerate(hist):
        nt = (h + spt - 1)//spt
        if nt == 0: continue
        c = 1 + (extra*nt)//T
        if c > nt: c = nt
        if cbase <= c_sel < cbase + c:
            j = c_sel - cbase; t0 = tbase + nt*j//c; t1 = tbase + nt*(j+1)//c; g_of = g
        if g == g_sel: c0, c1 = cbase, cbase + c
        cbase += c; tbase += nt
    return t0, t1, g_of, c0, c1



def test_scenario_partition_covers_every_tile_once():
    rng = np.random.default_rng(1)
    for trial in range(200):
        n_dom = int(rng.integers(1, 9))
        hist = [int(x) for x in rng.integers(0, 5000, size=n_dom) * (rng.random(n_dom) < 0.8)]
        B = sum(hist)
        if B == 0:
            continue
        spt = int(rng.integers(4, 129))
        G = min((B + spt - 1) // spt + n_dom, 296)
        tiles_key = []
        for g, h in enumerate(hist):
            tiles_key += [g] * ((h + spt - 1) // spt)
        covered = np.zeros(len(tiles_key), dtype=int)
        for c in range(G):
            t0, t1, g_of, _, _ = scenario_partition(hist, spt, G, c_sel=c)
            for t in range(t0, t1):
                covered[t] += 1
                assert tiles_key[t] == g_of
            if t1 > t0:
                _, _, _, c0, c1 = scenario_partition(hist, spt, G, g_sel=g_of)
                assert c0 <= c < c1
        assert (covered == 1).all()


RB, RADIX, TILE, THREADS, WARPS = 9, 512, 2048, 256, 8
def radix_pass(keys, vals, shift):
    N = len(keys); nblk = (N + TILE - 1)//TILE
    hist = np.zeros((RADIX, nblk), dtype=np.int64)
    for b in range(nblk):
        seg = keys[b*TILE:(b+1)*TILE]
        d = (seg >> shift) & (RADIX-1)
        np.add.at(hist[:, b], d, 1)
    flat = hist.reshape(-1)  # digit-major, block-minor
    # scan emul: 1024 threads, per = ceil(n/1024)
    n = flat.size; per = (n + 1023)//1024
    sums = np.array([flat[min(n, t*per):min(n, t*per+per)].sum() for t in range(1024)])
    excl = np.concatenate([[0], np.cumsum(sums)[:-1]])
    out = np.zeros_like(flat)
    for t in range(1024):
        a0 = min(n, t*per); a1 = min(n, a0+per); run = excl[t]
        for i in range(a0, a1):
            out[i] = run; run += flat[i]
    scanned = out.reshape(RADIX, nblk)
    ko = np.zeros_like(keys); vo = np.zeros_like(vals)
    for b in range(nblk):
        base = scanned[:, b].copy()
        cnt = np.zeros((WARPS, RADIX), dtype=np.int64)
        for j in range(0, TILE, THREADS):
            idx = b*TILE + j + np.arange(THREADS)
            valid = idx < N
            d = np.where(valid, (keys[np.minimum(idx, N-1)] >> shift) & (RADIX-1), -1)
            rank = np.zeros(THREADS, dtype=np.int64)
            for w in range(WARPS):
                lanes = slice(w*32, w*32+32)
                dd = d[lanes]
                for l in range(32):
                    rank[w*32+l] = np.sum(dd[:l] == dd[l])
                    if valid[w*32+l] and rank[w*32+l] == 0:
                        cnt[w, dd[l]] = np.sum(dd == dd[l])
            for t in range(THREADS):
                if valid[t]:
                    w = t//32
                    before = cnt[:w, d[t]].sum()
                    pos = base[d[t]] + before + rank[t]
                    ko[pos] = keys[idx[t]]; vo[pos] = vals[idx[t]]
            base += cnt.sum(axis=0); cnt[:] = 0
    return ko, vo


def test_radix_sort_is_stable():
    rng = np.random.default_rng(2)
    for trial in range(3):
        N = int(rng.integers(1, 5000))
        bits = int(rng.integers(1, 24))
        keys = rng.integers(0, 2 ** bits, size=N).astype(np.int64)
        if trial == 0:
            keys[:] = 5
        k, v = keys, np.arange(N)
        for p in range((bits + RB - 1) // RB):
            k, v = radix_pass(k, v, p * RB)
        order = np.argsort(keys, kind='stable')
        assert np.array_equal(k, keys[order]) and np.array_equal(v, order)


# ---------------------------------------------------------------------------------------------------------------------
# deferred table sweep (csrc/optim.cu: k_defer_begin / k_defer_catchup / k_adam_rows / k_defer_sweep, engine.train_step(defer=K))
# ---------------------------------------------------------------------------------------------------------------------
def _adam_row(p, m, v, g, c):
    """torch.optim.Adam's update of one row in float32 (the arithmetic both the model below and the reference loop use)."""
    f = np.float32
    m = (m + (g - m) * f(1 - c['b1'])).astype(f)
    v = (v * f(c['b2']) + f(1 - c['b2']) * g * g).astype(f)
    denom = (np.sqrt(v) / f(c['bc2_sqrt']) + f(c['eps'])).astype(f)
    p = (p - f(c['step_size']) * (m / denom)).astype(f)
    return p, m, v


def _consts(t, lr=0.005, b1=0.9, b2=0.999, eps=1e-8):
    return {'b1': b1, 'b2': b2, 'eps': eps, 'step_size': lr / (1 - b1 ** t), 'bc2_sqrt': np.sqrt(1 - b2 ** t)}


@pytest.mark.parametrize('K', [1, 4, 7, 16])
def test_deferred_sweep_model_equals_per_step_sweep(K):
    """CPU model of the deferred table sweep: every row that no batch read is regularised + Adam-updated every step in the
    reference (meta_basemodel.py:577-593, main.py:343).  The model below defers those updates exactly as the kernels do -- a ring
    of per-step constants, one byte per row counting the steps of the epoch already applied, catch-up of the rows a batch is
    about to read, a sweep that closes the epoch after K steps -- and must leave bit-identical rows, must hand the 'forward' rows
    that are up to date, and must report the same per-step sum of squares."""
    rng = np.random.default_rng(3)
    f = np.float32
    R, d, T, l2 = 60, 4, 23, f(1e-5)
    p0 = rng.normal(0, 1e-2, (R, d)).astype(f)
    touched = [np.unique(rng.integers(0, R, rng.integers(1, 12))) for _ in range(T)]
    grads = [rng.normal(0, 1e-3, (len(u), d)).astype(f) for u in touched]
    # --- the reference loop: every row every step
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    seen_ref, reg_ref = [], []
    for t in range(1, T + 1):
        seen_ref.append(p[touched[t - 1]].copy())                 # what the forward of step t reads
        reg_ref.append(float(np.sum(p.astype(np.float64) ** 2)))
        g = f(2) * l2 * p
        g[touched[t - 1]] += grads[t - 1]
        for r in range(R):
            p[r], m[r], v[r] = _adam_row(p[r], m[r], v[r], g[r], _consts(t))
    # --- the deferred model
    q, qm, qv = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    upto = np.zeros(R, dtype=np.uint8)
    ring, n_begun, reg_def = [], 0, {}
    reg_acc = {}

    def untouched_steps(r, j0, j1, base_step):
        for j in range(j0, j1):
            reg_acc[base_step + j] = reg_acc.get(base_step + j, 0.0) + float(np.sum(q[r].astype(np.float64) ** 2))
            q[r], qm[r], qv[r] = _adam_row(q[r], qm[r], qv[r], f(2) * l2 * q[r], ring[j])

    def sweep(base_step):
        nonlocal n_begun, ring
        for r in range(R):
            untouched_steps(r, int(upto[r]), n_begun, base_step)
            upto[r] = 0
        for j in range(n_begun):
            reg_def[base_step + j] = reg_acc.pop(base_step + j)
        n_begun, ring = 0, []

    base = 1                                                      # global index of the epoch's first step
    for t in range(1, T + 1):
        if n_begun >= K:                                          # the host closes the epoch before step t begins
            sweep(base)
            base = t
        ring.append(_consts(t)); n_begun += 1                     # sat_defer_begin
        cur = n_begun - 1
        rows = touched[t - 1]
        for r in rows:                                            # sat_defer_catchup
            untouched_steps(r, int(upto[r]), cur, base)
            upto[r] = cur
        assert np.array_equal(q[rows], seen_ref[t - 1]), t        # the forward gathers up-to-date rows
        for i, r in enumerate(rows):                              # k_adam_rows (deferred): this step's update, row marked current
            reg_acc[t] = reg_acc.get(t, 0.0) + float(np.sum(q[r].astype(np.float64) ** 2))
            q[r], qm[r], qv[r] = _adam_row(q[r], qm[r], qv[r], grads[t - 1][i] + f(2) * l2 * q[r], _consts(t))
            upto[r] = n_begun
        assert n_begun <= 16
    sweep(base)                                                   # sync_tables()
    assert np.array_equal(q, p) and np.array_equal(qm, m) and np.array_equal(qv, v)
    assert np.allclose([reg_def[t] for t in range(1, T + 1)], reg_ref, rtol=1e-12, atol=0)


# ---- index math of the ablation-flag kernels and of the per-scenario backward launches (round 2) ---------------------------------
def test_bilinear_gradient_is_the_diagonal_blocks_of_the_outer_product():
    """k_bwd_meta with SAT_FLAG_BILINEAR (backward.cu): the full 32x32 product M = qp^T dQ' is accumulated at sdTh[1024 + i*32 + j]
    by the shared 4x4-block routine; the gradient of theta (B_h[i][j] at theta[(h*dh + i)*dh + j], forward.cu) is then read out as
    sdTh[e] = sdTh[1024 + row*32 + (row/dh)*dh + j] with row = e / dh, j = e % dh.  Restated and checked against autograd-free
    algebra: d B_h = qp_h^T dQ'_h, and dqp_h = dQ'_h B_h^T."""
    rng = np.random.default_rng(0)
    D = 32
    for H in (1, 2, 4, 8):
        dh = D // H
        ntok = 57
        qp, dq = rng.standard_normal((ntok, D)), rng.standard_normal((ntok, D))
        theta = rng.standard_normal(D * dh)
        sd = np.zeros(4096)
        sd[1024:2048] = (qp.T @ dq).reshape(-1)                       # wgrad32(sA = qp, sG = dQ', acc = sdTh + 1024, ld = 32)
        for e in range(D * dh):                                       # the remap loop at the end of k_bwd_meta
            row, j = e // dh, e % dh
            sd[e] = sd[1024 + row * D + (row // dh) * dh + j]
        B = theta.reshape(H, dh, dh)
        for h in range(H):
            ref = qp[:, h * dh:(h + 1) * dh].T @ dq[:, h * dh:(h + 1) * dh]
            assert np.allclose(sd[h * dh * dh:(h + 1) * dh * dh].reshape(dh, dh), ref)
        # forward: q'_h[j] = sum_i qp_h[i] * theta[(h*dh + i)*dh + j];  backward (bilinear_dq): dqp_h[i] = sum_j dq'_h[j] * theta[(h*dh + i)*dh + j]
        fwd = np.stack([sum(qp[:, h * dh + i, None] * theta[(h * dh + i) * dh:(h * dh + i) * dh + dh][None, :] for i in range(dh))
                        for h in range(H)], axis=1).reshape(ntok, D)
        assert np.allclose(fwd, np.concatenate([qp[:, h * dh:(h + 1) * dh] @ B[h] for h in range(H)], axis=1))
        back = np.zeros((ntok, D))
        for h in range(H):
            for i in range(dh):
                back[:, h * dh + i] = dq[:, h * dh:(h + 1) * dh] @ theta[(h * dh + i) * dh:(h * dh + i) * dh + dh]
        assert np.allclose(back, np.concatenate([dq[:, h * dh:(h + 1) * dh] @ B[h].T for h in range(H)], axis=1))


def test_live_generated_weights_per_key():
    """SAT_META_PSZ (include/satrans_b200.h): 32 live generated weights per key with a gate, 1024 / H with the per-head bilinear map,
    4096 with the MetaNet -- the sizes of the reference's scenario-encoder output (satrans.py:156-162)."""
    GQ, GK, BIL, P = 128, 256, 512, 4096
    psz = lambda flags, H: 32 if flags & (GQ | GK) else (1024 // H if flags & BIL else P)
    for H in (1, 2, 4, 8):
        assert psz(GQ, H) == 32 and psz(GK | 1, H) == 32
        assert psz(BIL, H) == 32 * 32 // H
        assert psz(2 | 4 | 16, H) == 32 * 64 + 64 * 32


def test_per_scenario_sample_lists_cover_each_sample_once():
    """_ScenarioAttention.backward (selfatt.py): a stable sort of the clamped scenario ids gives contiguous, batch-ordered sample
    lists per scenario; samples outside 1..n are on no list (zero gradient)."""
    import torch
    g = torch.Generator().manual_seed(1)
    for n in (1, 3, 8):
        dom = torch.randint(-1, n + 3, (257,), generator=g, dtype=torch.int32)
        key = torch.where((dom >= 1) & (dom <= n), dom, torch.full_like(dom, n + 1))
        order = torch.argsort(key, stable=True).to(torch.int32)
        counts = torch.bincount(key.long(), minlength=n + 2)[1:n + 1].tolist()
        seen, start = [], 0
        for d, cnt in enumerate(counts):
            idx = order[start:start + cnt]
            assert bool((dom[idx.long()] == d + 1).all()) and bool((idx[1:] > idx[:-1]).all())      # one scenario, batch order
            seen += idx.tolist()
            start += cnt
        inside = ((dom >= 1) & (dom <= n)).nonzero().flatten().tolist()
        assert sorted(seen) == inside and len(seen) == len(set(seen))
