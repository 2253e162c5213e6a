"""The tridiagonal solve of the two-lanes-per-cell kernel, restated in numpy float32 and checked
against a float64 dense solve (CPU only; the kernel itself is covered by tests/test_gpu_pair.py).

hybrid9_b200/csrc/h9_physics_pair.cuh solves the 9-row system of HYDROLOGY.f90:755-843 (aquifer row 9
on top of the eight soil layers) from BOTH ends: the even lane eliminates rows 1..4 downwards, the odd
lane rows 9..5 upwards, and the two meet in a 2x2 system at the junction between rows 4 and 5; the
back substitution then runs outwards in both lanes at once.  Two formulations are checked:

* the one in the kernel: bet_n = dg_n - a_n c_{n-1} / bet_{n-1}, one reciprocal per row on the chain;
* the division-free one measured against the static schedule in profiles/r02/README.md section 8
  (N_n = dg_n N_{n-1} - a_n c_{n-1} N_{n-2}, U_n = r_n N_{n-1} - a_n U_{n-1}), kept here so that the
  next change to the kernel has its algebra pinned down first.
"""
import numpy as np
import pytest

F = np.float32


def systems(n_sys, seed, wet):
    """Rows of the size and dominance the sub-step produces: diagonal dz/dt + flux derivatives
    (dz 17.5 mm .. 1.5 m over 1800 s, the aquifer row up to 80 m), off-diagonals the flux
    derivatives, which grow with wetness but stay below the diagonal."""
    rng = np.random.default_rng(seed)
    dzdt = np.array([17.5, 27.6, 45.5, 75.0, 123.6, 203.8, 336.0, 553.9], np.float64) / 1800.0
    a = np.zeros((n_sys, 9))
    b = np.zeros((n_sys, 9))
    c = np.zeros((n_sys, 9))
    scale = 10.0 ** rng.uniform(-6, 0 if wet else -3, (n_sys, 9))
    lo = scale * rng.uniform(0.1, 1.0, (n_sys, 9))   # coupling to the row above
    up = scale * rng.uniform(0.1, 1.0, (n_sys, 9))   # coupling to the row below
    a[:, 1:] = -lo[:, 1:]
    c[:, :-1] = -up[:, :-1]
    b[:, :8] = dzdt + lo[:, :8] + up[:, :8]
    b[:, 8] = rng.uniform(1.0, 8e4, n_sys) / 1800.0 + lo[:, 8]
    r = rng.normal(0.0, 1e-4, (n_sys, 9))
    return a, b, c, r


def dense_solve(a, b, c, r):
    x = np.empty_like(r)
    for k in range(r.shape[0]):
        m = np.diag(b[k]) + np.diag(a[k, 1:], -1) + np.diag(c[k, :-1], 1)
        x[k] = np.linalg.solve(m, r[k])
    return x


def sides(a, b, c, r):
    """Each lane's rows in its own order, outer row first: (sub, diag, super, rhs) with `sub` the
    coupling to the row eliminated before and `super` the coupling to the next one."""
    even = (a[:, 0:4], b[:, 0:4], c[:, 0:4], r[:, 0:4])
    odd = (c[:, 8:3:-1], b[:, 8:3:-1], a[:, 8:3:-1], r[:, 8:3:-1])
    return even, odd


def sweep_reciprocal(side):
    sub, dg, sup, rhs = (v.astype(F) for v in side)
    n = dg.shape[1]
    gam = np.zeros_like(dg)
    u = np.zeros_like(dg)
    bet = dg[:, 0]
    u[:, 0] = rhs[:, 0] / bet
    gam[:, 0] = sup[:, 0] / bet
    for k in range(1, n):
        bet = dg[:, k] - sub[:, k] * gam[:, k - 1]
        rb = F(1.0) / bet
        u[:, k] = (rhs[:, k] - sub[:, k] * u[:, k - 1]) * rb
        gam[:, k] = sup[:, k] * rb
    return u, gam


def solve_reciprocal(a, b, c, r):
    even, odd = sides(a, b, c, r)
    (ue, ge), (uo, go) = sweep_reciprocal(even), sweep_reciprocal(odd)
    # junction: x4 + ge x5 = ue ; x5 + go x4 = uo
    den = F(1.0) - ge[:, -1] * go[:, -1]
    x4 = (ue[:, -1] - ge[:, -1] * uo[:, -1]) / den
    x5 = (uo[:, -1] - go[:, -1] * ue[:, -1]) / den
    x = np.zeros(r.shape, F)
    x[:, 3], x[:, 4] = x4, x5
    for k in range(2, -1, -1):
        x[:, k] = ue[:, k] - ge[:, k] * x[:, k + 1]
    for k in range(3, -1, -1):       # odd lane: local k <-> row 8 - k, its next row is 8 - k - 1
        x[:, 8 - k] = uo[:, k] - go[:, k] * x[:, 8 - k - 1]
    return x


def sweep_division_free(side):
    sub, dg, sup, rhs = (v.astype(F) for v in side)
    n = dg.shape[1]
    N = np.zeros_like(dg)
    U = np.zeros_like(dg)
    N[:, 0] = dg[:, 0]
    U[:, 0] = rhs[:, 0]
    for k in range(1, n):
        nm2 = N[:, k - 2] if k >= 2 else np.ones_like(dg[:, 0])
        N[:, k] = dg[:, k] * N[:, k - 1] - (sub[:, k] * sup[:, k - 1]) * nm2
        U[:, k] = rhs[:, k] * N[:, k - 1] - sub[:, k] * U[:, k - 1]
    return N, U, sup


def solve_division_free(a, b, c, r):
    even, odd = sides(a, b, c, r)
    (Ne, Ue, ce), (No, Uo, co) = sweep_division_free(even), sweep_division_free(odd)

    def g_last(N, sup):
        return sup[:, -1] * (N[:, -2])

    Ge, Go = g_last(Ne, ce), g_last(No, co)
    det = Ne[:, -1] * No[:, -1] - Ge * Go
    x4 = (Ue[:, -1] * No[:, -1] - Ge * Uo[:, -1]) / det
    x5 = (Uo[:, -1] * Ne[:, -1] - Go * Ue[:, -1]) / det
    x = np.zeros(r.shape, F)
    x[:, 3], x[:, 4] = x4, x5

    def back(N, U, sup, xs_next, k):
        nm1 = N[:, k - 1] if k >= 1 else np.ones_like(N[:, 0])
        rn = F(1.0) / N[:, k]
        return U[:, k] * rn - (sup[:, k] * nm1 * rn) * xs_next

    for k in range(2, -1, -1):
        x[:, k] = back(Ne, Ue, ce, x[:, k + 1], k)
    for k in range(3, -1, -1):
        x[:, 8 - k] = back(No, Uo, co, x[:, 8 - k - 1], k)
    return x, np.concatenate([Ne, No], axis=1)


@pytest.mark.parametrize("wet", [False, True])
def test_two_sided_solve_as_in_the_kernel(wet):
    a, b, c, r = systems(4000, 3, wet)
    ref = dense_solve(a, b, c, r)
    got = solve_reciprocal(a, b, c, r)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(got - ref) / scale).max() < 2e-5


@pytest.mark.parametrize("wet", [False, True])
def test_division_free_sweep_is_the_same_solve(wet):
    a, b, c, r = systems(4000, 5, wet)
    ref = dense_solve(a, b, c, r)
    got, N = solve_division_free(a, b, c, r)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(got - ref) / scale).max() < 2e-5
    # the running products stay far inside the float range (at most five pivots each)
    assert np.isfinite(N).all() and np.abs(N).min() > 1e-20 and np.abs(N).max() < 1e12
    # and the two formulations agree with each other at rounding level
    other = solve_reciprocal(a, b, c, r)
    assert (np.abs(got - other) / scale).max() < 2e-5
