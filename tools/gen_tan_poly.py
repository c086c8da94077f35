"""Generates the near-minimax polynomial coefficients of T(x) = tan(sqrt x)/sqrt x used by the sweep kernels
(rbqoc_b200/csrc/rbq_tanpoly.cuh): Chebyshev interpolation on [0, xmax] in 50-digit arithmetic, converted to the
monomial basis, rounded to binary64; prints the achieved relative error of the ROUNDED polynomial (exact arithmetic)."""
import mpmath as mp

mp.mp.dps = 60


def T(x):
    if x == 0:
        return mp.mpf(1)
    r = mp.sqrt(x)
    return mp.tan(r) / r


def cheb_monomial(xmax, deg):
    n = deg + 1
    nodes = [xmax / 2 * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * n))) for k in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x ** j
        b[i] = T(x)
    c = mp.lu_solve(A, b)
    return [c[j] for j in range(n)]


def max_rel_err(coefs, xmax, rounded=True, npts=4001):
    cs = [mp.mpf(float(c)) if rounded else c for c in coefs]
    worst = mp.mpf(0)
    for i in range(npts):
        x = xmax * i / (npts - 1)
        p = mp.mpf(0)
        for c in reversed(cs):
            p = p * x + c
        worst = max(worst, abs(p / T(x) - 1))
    return worst


if __name__ == "__main__":
    for name, xmax in (("A", mp.mpf("1.0e-3")), ("B", mp.mpf("4.0e-2"))):
        for deg in range(3, 10):
            c = cheb_monomial(xmax, deg)
            e = max_rel_err(c, xmax, rounded=False, npts=801)
            if e < mp.mpf("3e-18"):
                er = max_rel_err(c, xmax, rounded=True)
                print(f"// class {name}: x <= {mp.nstr(xmax, 3)}, degree {deg}, interpolation error {mp.nstr(e, 3)}, "
                      f"with binary64 coefficients {mp.nstr(er, 3)}")
                print("static constexpr double kTan%s[%d] = {%s};" % (name, deg + 1, ", ".join(repr(float(t)) for t in c)))
                break
