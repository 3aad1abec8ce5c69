"""The per-CTA tables of the superposition form are built by a blocked warp scan (mppi_tf_b200/csrc/mppi_linear.cuh,
build_linear_tables): lane l runs its chunk of ceil(T/32) steps from a zero state, the chunk responses are composed as affine
maps, and the lane replays its chunk from its true incoming state.  This restates that scheme in numpy (fp64) and checks it
against the serial recurrences it replaces - the model step of /root/reference/src/model_base.cpp:53-82 forward and its
adjoint backward - for horizons that do and do not fill the 32 lanes.  CPU only: the algebra, not the kernel."""
import numpy as np
import pytest


def serial(U, p0, v0, gp, gv, dt, c_pu, c_vu, sqp, sqv, a1, b1, b2, lin):
    T = len(U)
    p, v = p0, v0
    dp, dv = np.zeros(T), np.zeros(T)
    for t in range(T):
        p = p + dt * v + c_pu * U[t]
        v = v + c_vu * U[t]
        dp[t], dv[t] = sqp * (p - gp), sqv * (v - gv)
    aP = aV = 0.0
    L = np.zeros(T)
    for t in range(T, 0, -1):
        w2 = 4.0 if t == T else 2.0
        aV = aV + a1 * aP + w2 * dv[t - 1]
        aP = aP + w2 * dp[t - 1]
        L[t - 1] = b1 * aP + b2 * aV + lin[t - 1]
    return dp, dv, L


def blocked_scan(U, p0, v0, gp, gv, dt, c_pu, c_vu, sqp, sqv, a1, b1, b2, lin, lanes=32):
    T = len(U)
    Lc = (T + lanes - 1) // lanes
    lo = [min(T, l * Lc) for l in range(lanes)]
    hi = [min(T, lo[l] + Lc) for l in range(lanes)]
    # forward: zero-state chunk responses, inclusive Kogge-Stone scan of (p, v, n), replay
    sp, sv, sn = np.zeros(lanes), np.zeros(lanes), np.zeros(lanes, int)
    for l in range(lanes):
        bp = bv = 0.0
        for t in range(lo[l], hi[l]):
            bp = bp + dt * bv + c_pu * U[t]
            bv = bv + c_vu * U[t]
        sp[l], sv[l], sn[l] = bp, bv, hi[l] - lo[l]
    o = 1
    while o < lanes:
        lp, lv, ln = np.roll(sp, o), np.roll(sv, o), np.roll(sn, o)
        for l in range(lanes - 1, o - 1, -1):
            sp[l] = sp[l] + lp[l] + sn[l] * dt * lv[l]
            sv[l] = sv[l] + lv[l]
            sn[l] = sn[l] + ln[l]
        o *= 2
    dp, dv = np.zeros(T), np.zeros(T)
    for l in range(lanes):
        ip, iv = (sp[l - 1], sv[l - 1]) if l else (0.0, 0.0)
        p, v = p0 + lo[l] * dt * v0 + ip, v0 + iv
        for t in range(lo[l], hi[l]):
            p = p + dt * v + c_pu * U[t]
            v = v + c_vu * U[t]
            dp[t], dv[t] = sqp * (p - gp), sqv * (v - gv)
    # backward: lane l walks steps t in (u_lo, u_hi] downwards from T - l Lc
    uh = [max(0, T - l * Lc) for l in range(lanes)]
    ul = [max(0, uh[l] - Lc) for l in range(lanes)]
    aPs, aVs, an = np.zeros(lanes), np.zeros(lanes), np.zeros(lanes, int)
    for l in range(lanes):
        cP = cV = 0.0
        for t in range(uh[l], ul[l], -1):
            w2 = 4.0 if t == T else 2.0
            cV = cV + a1 * cP + w2 * dv[t - 1]
            cP = cP + w2 * dp[t - 1]
        aPs[l], aVs[l], an[l] = cP, cV, uh[l] - ul[l]
    o = 1
    while o < lanes:
        lP, lV, ln = np.roll(aPs, o), np.roll(aVs, o), np.roll(an, o)
        for l in range(lanes - 1, o - 1, -1):
            aVs[l] = aVs[l] + lV[l] + an[l] * a1 * lP[l]
            aPs[l] = aPs[l] + lP[l]
            an[l] = an[l] + ln[l]
        o *= 2
    L = np.zeros(T)
    for l in range(lanes):
        aP, aV = (aPs[l - 1], aVs[l - 1]) if l else (0.0, 0.0)
        for t in range(uh[l], ul[l], -1):
            w2 = 4.0 if t == T else 2.0
            aV = aV + a1 * aP + w2 * dv[t - 1]
            aP = aP + w2 * dp[t - 1]
            L[t - 1] = b1 * aP + b2 * aV + lin[t - 1]
    return dp, dv, L


@pytest.mark.parametrize("T", [1, 2, 7, 20, 31, 32, 33, 50, 100, 257, 1024])
def test_blocked_scan_equals_the_serial_recurrences(T):
    rng = np.random.default_rng(T)
    U = rng.standard_normal(T)
    lin = rng.standard_normal(T)
    args = dict(p0=0.3, v0=-0.7, gp=1.0, gv=0.2, dt=0.1, c_pu=0.005 / 1.5, c_vu=0.1 / 1.5, sqp=1.3, sqv=0.8, a1=0.1 * 1.3 / 0.8,
                b1=0.4, b2=0.9, lin=lin)
    want = serial(U, **args)
    got = blocked_scan(U, **args)
    for g, w, name in zip(got, want, ("D_p", "D_v", "L")):
        assert np.abs(g - w).max() <= 1e-12 * max(1.0, np.abs(w).max()), name
