"""numpy emulation of the pipeline logic of k_stream3d (mg_stream3d.cuh): same step/stage
timing, ring parity, active/emit predicates and masks, vectorised over the tile. Used to
validate the index arithmetic on the CPU before spending GPU time. Not product code."""
import sys
import numpy as np


def jacobi_ref(u, f, h):
    p = np.pad(u, 1)
    S = ((((p[1:-1, 1:-1, :-2] + p[1:-1, 1:-1, 2:]) + p[1:-1, :-2, 1:-1]) + p[1:-1, 2:, 1:-1]) + p[:-2, 1:-1, 1:-1]) + p[2:, 1:-1, 1:-1]
    return (f - S / h**2) / (-6 / h**2)


def residual_ref(u, f, h):
    p = np.pad(u, 1)
    S = ((((p[1:-1, 1:-1, :-2] + p[1:-1, 1:-1, 2:]) + p[1:-1, :-2, 1:-1]) + p[1:-1, 2:, 1:-1]) + p[:-2, 1:-1, 1:-1]) + p[2:, 1:-1, 1:-1]
    return f - (S / h**2 + (-6 / h**2) * u)


def restrict_ref(r):
    s = r[0::2, 0::2, 0::2] + r[0::2, 0::2, 1::2]
    s = s + r[0::2, 1::2, 0::2]
    s = s + r[0::2, 1::2, 1::2]
    s = s + r[1::2, 0::2, 0::2]
    s = s + r[1::2, 0::2, 1::2]
    s = s + r[1::2, 1::2, 0::2]
    s = s + r[1::2, 1::2, 1::2]
    return .125 * s


def emulate(src, f, h, S, RES, PRO, Vp, TX, TY, TZ, VX=4, NSLOT=4):
    L = src.shape[0]
    NST = S + (1 if RES else 0)
    H = NST
    HX = (H + VX - 1) // VX * VX
    HY = (H + 1) // 2 * 2 if RES else H
    WX, WY = TX + 2 * HX, TY + 2 * HY
    dst = np.full_like(src, np.nan)
    Rout = np.full((L // 2,) * 3, np.nan)
    inv_h2, adiag = 1 / h**2, -6 / h**2

    def load_plane(x0, y0, z):  # TMA with zero fill
        out = np.zeros((WY, WX))
        if 0 <= z < L:
            for yy in range(WY):
                gy = y0 - HY + yy
                if 0 <= gy < L:
                    xs = np.arange(WX) + x0 - HX
                    ok = (xs >= 0) & (xs < L)
                    out[yy, ok] = src[z, gy, xs[ok]]
        return out

    for bz in range((L + TZ - 1) // TZ):
        for by in range((L + TY - 1) // TY):
            for bx in range((L + TX - 1) // TX):
                x0, y0, z0 = bx * TX, by * TY, bz * TZ
                z1 = min(z0 + TZ, L)
                tz = z1 - z0
                zb, nin, T = z0 - H, tz + 2 * H, tz + 3 * H - 1
                gx = np.arange(WX) + x0 - HX
                gy = np.arange(WY) + y0 - HY
                inm = ((gy >= 0) & (gy < L))[:, None] & ((gx >= 0) & (gx < L))[None, :]
                stm = inm & ((np.arange(WY) >= HY) & (np.arange(WY) < HY + TY))[:, None] & \
                    ((np.arange(WX) >= HX) & (np.arange(WX) < HX + TX))[None, :]
                slots = [None] * NSLOT
                ring = [[np.full((WY, WX), np.nan), np.full((WY, WX), np.nan)] for _ in range(NST)]
                acc = [np.zeros((WY, WX)) for _ in range(NST)]
                prev = [np.zeros((WY, WX)) for _ in range(NST)]
                rpart = np.zeros((WY // 2, WX // 2))

                def fix(t):
                    p = zb + t
                    if not PRO or p < 0 or p >= L:
                        return
                    sl = slots[t % NSLOT]
                    for yy in range(WY):
                        for xx in range(WX):
                            if inm[yy, xx]:
                                sl[yy, xx] = sl[yy, xx] + Vp[p >> 1, gy[yy] >> 1, gx[xx] >> 1]

                for k in range(NSLOT - 1):
                    if k < nin:
                        slots[k] = load_plane(x0, y0, zb + k)
                if PRO:
                    fix(0)
                for t in range(T):
                    k = t + NSLOT - 1
                    if k < nin:
                        slots[k % NSLOT] = load_plane(x0, y0, zb + k)
                    writes = []
                    for sidx in range(NST):
                        s = sidx + 1
                        active = (t >= 3 * sidx) and (t <= nin + s - 2)
                        if not active:
                            continue
                        emit = t >= 3 * s - 1
                        p = zb + t - 2 * sidx - 1
                        inp = slots[t % NSLOT] if sidx == 0 else ring[sidx - 1][(t - 1) & 1]
                        is_res = RES and s == NST
                        pin = 0 <= p < L
                        fv = np.zeros((WY, WX))
                        if emit and pin:
                            for yy in range(WY):
                                if 0 <= gy[yy] < L:
                                    ok = (gx >= 0) & (gx < L)
                                    fv[yy, ok] = f[p, gy[yy], gx[ok]]
                        c = inp
                        # clamped neighbours (garbage zone at the tile edge)
                        xl = np.concatenate([c[:, :1], c[:, :-1]], 1)
                        xr = np.concatenate([c[:, 1:], c[:, -1:]], 1)
                        yl = np.concatenate([c[:1, :], c[:-1, :]], 0)
                        yr = np.concatenate([c[1:, :], c[-1:, :]], 0)
                        with np.errstate(invalid="ignore"):
                            part = ((xl + xr) + yl) + yr
                            tot = acc[sidx] + c
                            if is_res:
                                o = fv - (tot * inv_h2 + adiag * prev[sidx])
                            else:
                                o = (fv - tot * inv_h2) / adiag
                            outv = np.where(inm & pin, o, 0.0)
                            acc[sidx] = part + prev[sidx]
                        prev[sidx] = c.copy()
                        if not emit:
                            continue
                        if not is_res:
                            if s < NST:
                                writes.append((sidx, t & 1, outv))
                            if s == S and z0 <= p < z1:
                                for yy in range(WY):
                                    for xx in range(WX):
                                        if stm[yy, xx]:
                                            dst[p, gy[yy], gx[xx]] = outv[yy, xx]
                        elif z0 <= p < z1:
                            a_ = outv[0::2, 0::2] + outv[0::2, 1::2]
                            a_ = a_ + outv[1::2, 0::2]
                            if p % 2 == 0:
                                rpart = a_ + outv[1::2, 1::2]
                            else:
                                sacc = rpart + outv[0::2, 0::2]
                                sacc = sacc + outv[0::2, 1::2]
                                sacc = sacc + outv[1::2, 0::2]
                                sacc = sacc + outv[1::2, 1::2]
                                for cy in range(WY // 2):
                                    for cx in range(WX // 2):
                                        if stm[2 * cy, 2 * cx] and stm[2 * cy + 1, 2 * cx + 1]:
                                            Rout[p >> 1, gy[2 * cy] >> 1, gx[2 * cx] >> 1] = .125 * sacc[cy, cx]
                    for sidx, par, o in writes:   # stage outputs become visible after the barrier
                        ring[sidx][par] = o
                    if PRO and t + 1 < nin:
                        fix(t + 1)
    return dst, Rout


def main():
    rng = np.random.default_rng(0)
    ok = True
    for L, TX, TY, TZ in ((16, 8, 4, 8), (16, 8, 8, 16), (32, 16, 8, 6)):
        h = 1.0 / L
        src = rng.uniform(-1, 1, (L, L, L))
        f = rng.uniform(-1, 1, (L, L, L)) * L * L
        Vp = rng.uniform(-1, 1, (L // 2,) * 3)
        for S in (1, 2, 3, 4):
            for RES, PRO in ((False, False), (True, False), (False, True)):
                if RES and S == 4:
                    continue
                u = src.copy()
                if PRO:
                    u = u + np.repeat(np.repeat(np.repeat(Vp, 2, 0), 2, 1), 2, 2)
                for _ in range(S):
                    u = jacobi_ref(u, f, h)
                dst, Rout = emulate(src, f, h, S, RES, PRO, Vp, TX, TY, TZ)
                good = np.array_equal(dst, u)
                if RES:
                    good = good and np.array_equal(Rout, restrict_ref(residual_ref(u, f, h)))
                print(f"L={L} tile={TX}x{TY}x{TZ} S={S} RES={RES} PRO={PRO}: {'ok' if good else 'MISMATCH'}")
                ok = ok and good
    sys.exit(0 if ok else 1)


if __name__ == "__main__" and "--rings" not in sys.argv:
    main()


def check_rings(NST, TZ, NSLOT=3):
    """Slot/phase bookkeeping of the source ring and the f ring exactly as k_stream3d does it: cursors su/pu and sf;
    at step t the TMA thread issues source plane t+2 AND f plane t+1 on ONE mbarrier (the source slot's, expect_tx =
    two planes); the prologue issues source planes 0, 1 and puts f plane 0 on the mbarrier of source plane 1. Checked:
    a slot is never refilled while a stage still needs its tenant, the parity every wait uses is the phase that
    carries the awaited planes, and the single wait of step t (source slot of plane t) also covers the f plane
    stage 1 reads at step t (plane t-1)."""
    assert NSLOT == 3
    H = NST
    NF = 2 * NST + 1
    nin, T = TZ + 2 * H, TZ + 3 * H - 1
    uslot, uphase = [None] * NSLOT, [0] * NSLOT      # tenant plane; completed phases of the slot's mbarrier
    bar_f = [None] * NSLOT                          # f plane counted on the current phase of each mbarrier
    fslot = [None] * NF
    f_last_use = {}
    f_waited = set()                                # f planes whose mbarrier phase some step has awaited
    for k in range(NSLOT - 1):                      # prologue
        uslot[k] = k; uphase[k] += 1
    fslot[0] = 0; bar_f[1] = 0
    su, pu, sf = 0, 0, NF - 1
    for t in range(T):
        k = t + NSLOT - 1
        if k < nin:
            ks = NSLOT - 1 if su == 0 else su - 1
            assert ks == k % NSLOT
            assert uslot[ks] is None or uslot[ks] <= t - 1, "source slot still in use"
            j = t + 1
            assert j <= nin - 2                      # the two issue conditions are the same condition
            ksf = (sf + 2) % NF
            assert ksf == j % NF
            old = fslot[ksf]
            assert old is None or f_last_use.get(old, -1) < t, ("f slot still in use", t, old)
            uslot[ks] = k; uphase[ks] += 1; bar_f[ks] = j; fslot[ksf] = j
        else:
            assert t + 1 > nin - 2
        if t < nin:                                  # the step's single wait
            assert uslot[su] == t and (uphase[su] - 1) & 1 == pu, ("source phase", t)
            if bar_f[su] is not None:
                f_waited.add(bar_f[su])
        for sidx in range(NST):
            s = sidx + 1
            if not (t >= 3 * sidx and t <= nin + s - 2):
                continue
            if t >= 3 * s - 1:
                jj = t - 1 - 2 * sidx
                kf = (sf - 2 * sidx) % NF
                assert 1 <= jj <= nin - 2 and fslot[kf] == jj, ("f plane", t, s, jj, fslot[kf])
                assert jj in f_waited, ("f plane read before its mbarrier was awaited", t, s, jj)
                f_last_use[jj] = t
        su += 1
        if su == NSLOT:
            su, pu = 0, pu ^ 1
        sf += 1
        if sf == NF:
            sf = 0
    assert 0 in f_waited                            # the never-used f plane 0 has landed before the CTA can exit
    return True


if __name__ == "__main__" and "--rings" in sys.argv:
    for NST in (1, 2, 3, 4, 5):
        for TZ in (2, 4, 6, 16, 33, 64):
            check_rings(NST, TZ)
    print("ring bookkeeping ok")
