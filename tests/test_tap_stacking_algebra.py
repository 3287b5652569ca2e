"""Round-2 groundwork (CPU only): the algebra of *tap stacking along N* for the tensor-core
convolution.  A numpy model of the implicit GEMM on the zero-haloed packed
grid shows that stacking the S weight taps (dy, dz=0..S-1) of one (dx, dy) into ONE B operand of
N = S*Cout columns, multiplying it with the single A view shifted by off(dy, dz=0), and folding the
column blocks back with a row shift  out[q] = sum_c D_c[q + c]  reproduces the 3x3x3 convolution,
provided row tiles advance by 128 - (S-1) rows (the last S-1 rows of a tile are incomplete and are
recomputed as the first rows of the next tile)."""
import numpy as np
import pytest


def packed_grid(x):                    # x [C, r, r, r] -> rows [(r+2)^3, C] with a zero halo
    C, r = x.shape[0], x.shape[1]
    g = np.zeros((r + 2, r + 2, r + 2, C), np.float64)
    g[1:-1, 1:-1, 1:-1] = np.transpose(x, (1, 2, 3, 0))
    return g.reshape(-1, C)


def conv_reference(x, w):              # direct 3x3x3, padding 1: x [C,r,r,r], w [O,C,3,3,3] -> [O,r,r,r]
    C, r = x.shape[0], x.shape[1]
    xp = np.zeros((C, r + 2, r + 2, r + 2))
    xp[:, 1:-1, 1:-1, 1:-1] = x
    out = np.zeros((w.shape[0], r, r, r))
    for dx in range(3):
        for dy in range(3):
            for dz in range(3):
                out += np.einsum('oc,cxyz->oxyz', w[:, :, dx, dy, dz], xp[:, dx:dx + r, dy:dy + r, dz:dz + r])
    return out


@pytest.mark.parametrize("S", [2, 3])
@pytest.mark.parametrize("r,C,O", [(6, 8, 4), (9, 4, 8)])
def test_stacked_taps_reproduce_the_convolution(S, r, C, O):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((C, r, r, r))
    w = rng.standard_normal((O, C, 3, 3, 3))
    rows = packed_grid(x)
    rp = r + 2
    P = rp ** 3
    guard = 128 + rp * rp + rp + 8            # (the kernel reads up to one tile past p_end: masked rows)
    buf = np.zeros((P + 2 * guard, C))
    buf[guard:guard + P] = rows                                   # guard rows like alloc_vg()
    off = lambda dx, dy, dz: (dx - 1) * rp * rp + (dy - 1) * rp + (dz - 1)
    p_begin, p_end = rp * rp, (rp - 1) * rp * rp                  # geom_grid(): interior x planes
    M, pitch = 128, 128 - (S - 1)
    out_rows = np.zeros((P, O))
    ntile = -(-(p_end - p_begin) // pitch)
    for t in range(ntile):
        p0 = p_begin + t * pitch
        D = np.zeros((M, 3 * O))                                  # accumulator: column block c <-> tap dz = c
        for dx in range(3):
            for dy in range(3):
                # stacked MMA over dz = 0..S-1: ONE A view, shifted by the offset of the first tap of the stack
                a = buf[guard + p0 + off(dx, dy, 0): guard + p0 + off(dx, dy, 0) + M]
                for c in range(S):
                    D[:, c * O:(c + 1) * O] += a @ w[:, :, dx, dy, c].T
                # taps that did not fit the stack (S = 2: dz = 2) run as single MMAs on block 0's alignment
                for dz in range(S, 3):
                    a1 = buf[guard + p0 + off(dx, dy, dz): guard + p0 + off(dx, dy, dz) + M]
                    D[:, 0:O] += a1 @ w[:, :, dx, dy, dz].T
        # epilogue: out[q] = D_0[q] + D_1[q+1] (+ D_2[q+2]); lanes >= pitch are incomplete -> next tile
        for lane in range(pitch):
            q = p0 + lane
            if q >= p_end:
                break
            v = D[lane, 0:O].copy()
            for c in range(1, S):
                v += D[lane + c, c * O:(c + 1) * O]
            out_rows[q] = v
    got = out_rows.reshape(rp, rp, rp, O)[1:-1, 1:-1, 1:-1]
    ref = np.transpose(conv_reference(x, w), (1, 2, 3, 0))
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("S,NT,KG,cin_pad,cout_pad", [(3, 64, 8, 64, 64), (3, 32, 8, 36, 32), (2, 128, 4, 128, 256)])
def test_stacked_weight_packing_matches_the_issuers_addressing(S, NT, KG, cin_pad, cout_pad):
    """Python transcription of tc::k_pack_tc_stack's index decode and of the B-operand addresses the
    UMMA issuers of tc::k_conv_stack form (16-byte units: one unit = 4 input channels of one output
    channel): every element the tensor core would read must be the weight of the tap / channel pair
    the algebra above assumes."""
    rng = np.random.default_rng(1)
    nchunk = -(-(cin_pad // 4) // KG)
    wt = rng.standard_normal((27, cin_pad, cout_pad))            # SIMT packing wt[tap][ci][co]
    per_dy = 3 * KG * NT
    n_nt = cout_pad // NT
    total = n_nt * nchunk * 3 * 3 * per_dy * 4
    w = np.zeros(total)
    for i in range(total):                                       # == k_pack_tc_stack
        jj = i % 4
        r = i // 4
        e = r % per_dy; r //= per_dy
        dy = r % 3; r //= 3
        dx = r % 3; r //= 3
        cc = r % nchunk; r //= nchunk
        nt = r
        if e < KG * S * NT:
            kg, n2 = divmod(e, S * NT)
            dz, n = divmod(n2, NT)
        else:
            e2 = e - KG * S * NT
            dz = S + e2 // (KG * NT)
            kg, n = divmod(e2 % (KG * NT), NT)
        tap = (dx * 3 + dy) * 3 + dz
        ci = (cc * KG + kg) * 4 + jj
        w[i] = wt[tap, ci, nt * NT + n] if ci < cin_pad else 0.0
    stage_units = 9 * KG * NT                                    # one (chunk, dx) stage, in 16-byte units
    SN = S * NT
    for nt in range(n_nt):
        for cc in range(nchunk):
            for dx in range(3):
                base = ((nt * nchunk + cc) * 3 + dx) * stage_units   # producer: wsrc + (cc*3 + tg) * stage
                for dy in range(3):
                    b_dy = base + dy * per_dy
                    for k2 in range(0, KG, 2):
                        for k in range(8):                          # K = 8 of one UMMA: two k-groups
                            kg = k2 + k // 4
                            ci = (cc * KG + kg) * 4 + k % 4
                            for n2 in (0, 1, NT - 1, NT, SN - 1):   # stacked MMA: start b_dy + k2*SN, LBO = SN
                                unit = b_dy + kg * SN + n2
                                want = wt[(dx * 3 + dy) * 3 + n2 // NT, ci, nt * NT + n2 % NT] if ci < cin_pad else 0.0
                                assert w[unit * 4 + k % 4] == want
                            for dz in range(S, 3):                  # single taps: start ... + KG*SN + (dz-S)*KG*NT, LBO = NT
                                for n in (0, NT - 1):
                                    unit = b_dy + KG * SN + (dz - S) * KG * NT + kg * NT + n
                                    want = wt[(dx * 3 + dy) * 3 + dz, ci, nt * NT + n] if ci < cin_pad else 0.0
                                    assert w[unit * 4 + k % 4] == want


@pytest.mark.parametrize("S", [2, 3])
def test_epilogue_fold_via_shuffles_and_quarter_exchange(S):
    """Transcription of the epilogue of tc::k_conv_stack for one 16-column chunk: four warps (TMEM lane
    quarters) of 32 lanes; lane l of quarter q folds D_c from lane l+c -- __shfl_down inside the warp, the
    rows that cross a quarter come from the next quarter through the shared-memory exchange (slot 0 =
    block 1 / lane 0, slots 1, 2 = block 2 / lanes 0, 1)."""
    rng = np.random.default_rng(2)
    D = rng.standard_normal((S, 128, 16))                         # [column block][tile row][column]
    TP = 128 - (S - 1)
    xch = np.full((4, 3, 16), np.nan)
    for q in range(4):                                            # writers (before the barrier)
        for c in range(1, S):
            for lane in range(32):
                if q > 0 and lane < c:
                    slot = 0 if c == 1 else 1 + lane
                    xch[q, slot] = D[c, q * 32 + lane]
    for q in range(4):                                            # readers
        for lane in range(32):
            acc = D[0, q * 32 + lane].copy()
            for c in range(1, S):
                if lane + c < 32:
                    sh = D[c, q * 32 + lane + c]                  # __shfl_down_sync(v, c)
                else:
                    slot = 0 if c == 1 else 1 + (lane + c - 32)
                    sh = xch[q + 1, slot] if q < 3 else np.zeros(16)
                acc = acc + sh
            lt = q * 32 + lane
            if lt < TP:                                           # rows >= TP are recomputed by the next tile
                want = sum(D[c, lt + c] for c in range(S))
                assert np.allclose(acc, want)
