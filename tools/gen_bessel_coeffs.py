#!/usr/bin/env python
"""Generate the polynomial coefficients used for K0, K1 (fluxPoint) and J1
(fluxFormIntegrand) by both the CUDA path and the CPU oracle.

The reference calls GSL's gsl_sf_bessel_K0/K1/J1 (src/UpcCrossSection.cpp:171-172,189).
GSL is not in this image, so instead of recalling GSL's Chebyshev tables we derive our own
near-minimax (Chebyshev-interpolant) fits with mpmath at 50 digits; the fits are accurate to
a few 1e-17 and evaluated by Horner in a centred variable u in [-1,1].

Forms (u is always the centred variable of the stated range):
  x <= 2 :  K0(x) = -ln(x/2) I0(x) + P0(y),      I0(x) = Q0(y),    y = x^2 in [0,4]
            K1(x) =  ln(x/2) I1(x) + P1(y)/x,    I1(x) = x Q1(y)
  2<x<=8 :  Kn(x) = exp(-x)/sqrt(x) * A_n(u),    u = (16/x - 5)/3
  x > 8  :  Kn(x) = exp(-x)/sqrt(x) * B_n(u),    u = 16/x - 1
  x <= 8 :  J1(x) = x * PJ(u),                   u = x^2/32 - 1
  x > 8  :  J1(x) = sqrt(2/(pi x)) * M(w) * sin(x - pi/4 + T(w)/x),  w = 64/x^2, u = 2w-1
            (modulus/phase form; sin(x-pi/4+eps) is expanded as GSL does so that x itself is
             the only large argument handed to sincos)

Writes one header with plain `static const double` arrays; --out selects the path.  The same
script writes the product copy (upcgen_b200/csrc/upc_bessel_coeffs.h) and the oracle copy
(oracle/upc_bessel_coeffs.h) so that the oracle does not include product code.
"""
import argparse
import mpmath as mp

mp.mp.dps = 50


def cheb_coeffs(f, n):
    """Chebyshev coefficients c_0..c_{n-1} of the degree n-1 interpolant of f on [-1,1]."""
    xs = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    fs = [f(x) for x in xs]
    cs = []
    for j in range(n):
        s = mp.fsum(fs[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n))
        cs.append(2 * s / n)
    cs[0] /= 2
    return cs


def cheb_to_mono(cs):
    """Convert a Chebyshev series to monomial coefficients (exact in mp arithmetic)."""
    n = len(cs)
    # T_0 = 1, T_1 = x, T_{k+1} = 2x T_k - T_{k-1}
    T_prev = [mp.mpf(1)]
    T_cur = [mp.mpf(0), mp.mpf(1)]
    out = [mp.mpf(0)] * n
    out[0] += cs[0]
    if n > 1:
        out[1] += cs[1]
    for k in range(2, n):
        T_next = [mp.mpf(0)] * (k + 1)
        for i, a in enumerate(T_cur):
            T_next[i + 1] += 2 * a
        for i, a in enumerate(T_prev):
            T_next[i] -= a
        for i, a in enumerate(T_next):
            out[i] += cs[k] * a
        T_prev, T_cur = T_cur, T_next
    return out


def fit(f, tol, nmax=40):
    """Smallest n whose trailing Chebyshev coefficients fall below tol (relative to c_0)."""
    for n in range(6, nmax):
        cs = cheb_coeffs(f, n + 4)
        scale = max(abs(c) for c in cs)
        if all(abs(c) < tol * scale for c in cs[n:]):
            cs = cheb_coeffs(f, n)
            return cheb_to_mono(cs)
    raise RuntimeError("fit did not converge")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--device", action="store_true", help="emit the CUDA copy (__constant__ arrays)")
    args = ap.parse_args()
    tol = mp.mpf("2e-18")

    def y_of(u):  # y = x^2 in [0,4]
        return 2 * (u + 1)

    def P0(u):
        y = y_of(u)
        if y == 0:
            return -mp.euler
        x = mp.sqrt(y)
        return mp.besselk(0, x) + mp.log(x / 2) * mp.besseli(0, x)

    def Q0(u):
        return mp.besseli(0, mp.sqrt(y_of(u)))

    def P1(u):
        y = y_of(u)
        if y == 0:
            return mp.mpf(1)
        x = mp.sqrt(y)
        return x * (mp.besselk(1, x) - mp.log(x / 2) * mp.besseli(1, x))

    def Q1(u):
        y = y_of(u)
        if y == 0:
            return mp.mpf(1) / 2
        x = mp.sqrt(y)
        return mp.besseli(1, x) / x

    def KA(nu):
        def f(u):
            x = 16 / (3 * u + 5)
            return mp.besselk(nu, x) * mp.exp(x) * mp.sqrt(x)
        return f

    def KB(nu):
        def f(u):
            t = (u + 1) / 16  # 1/x
            if t == 0:
                return mp.sqrt(mp.pi / 2)
            x = 1 / t
            return mp.besselk(nu, x) * mp.exp(x) * mp.sqrt(x)
        return f

    def PJ(u):
        y = 32 * (u + 1)
        if y == 0:
            return mp.mpf(1) / 2
        x = mp.sqrt(y)
        return mp.besselj(1, x) / x

    def modphase(x):
        j = mp.besselj(1, x)
        yv = mp.bessely(1, x)
        M = mp.sqrt(j * j + yv * yv)
        # J1 = M cos(theta), Y1 = M sin(theta); theta = x - 3pi/4 + delta
        theta = mp.atan2(yv, j)
        delta = theta - (x - 3 * mp.pi / 4)
        # bring delta to the principal branch near 3/(8x)
        delta = delta - 2 * mp.pi * mp.nint(delta / (2 * mp.pi))
        return M, delta

    def JM(u):
        w = (u + 1) / 2
        if w == 0:
            return mp.mpf(1)
        x = 8 / mp.sqrt(w)
        M, _ = modphase(x)
        return M * mp.sqrt(mp.pi * x / 2)

    def JT(u):
        w = (u + 1) / 2
        if w == 0:
            return mp.mpf(3) / 8
        x = 8 / mp.sqrt(w)
        _, d = modphase(x)
        return d * x

    tables = [
        ("K0_P", P0), ("K0_Q", Q0), ("K1_P", P1), ("K1_Q", Q1),
        ("K0_A", KA(0)), ("K1_A", KA(1)), ("K0_B", KB(0)), ("K1_B", KB(1)),
        ("J1_P", PJ), ("J1_M", JM), ("J1_T", JT),
    ]
    lines = [
        "/* GENERATED by tools/gen_bessel_coeffs.py -- do not edit.",
        " * Monomial coefficients (ascending powers of the centred variable u in [-1,1]) of",
        " * Chebyshev-interpolant fits computed with mpmath at 50 digits.  See the generator",
        " * for the functional forms.  Replaces gsl_sf_bessel_K0/K1/J1 (reference",
        " * src/UpcCrossSection.cpp:171-172,189), which are not available in this image. */",
        "#pragma once",
        "#ifndef UPC_BESSEL_CONST",
        "#define UPC_BESSEL_CONST " + ("static __constant__ double" if args.device else "static const double"),
        "#endif",
    ]
    for name, f in tables:
        co = fit(f, tol)
        lines.append(f"#define UPC_{name}_N {len(co)}")
        if args.device:
            # value list as a macro too, so that several tables can be concatenated into the one
            # __constant__ block the QAGS kernel streams with LDCU (upc_hot.cuh)
            lines.append(f"#define UPC_{name}_VALUES \\")
            for i, c in enumerate(co):
                lines.append("  " + mp.nstr(c, 20, min_fixed=0, max_fixed=0) + ("" if i == len(co) - 1 else ", \\"))
            lines.append(f"UPC_BESSEL_CONST UPC_{name}[{len(co)}] = {{UPC_{name}_VALUES}};")
        else:
            lines.append(f"UPC_BESSEL_CONST UPC_{name}[{len(co)}] = {{")
            for c in co:
                lines.append("  " + mp.nstr(c, 20, min_fixed=0, max_fixed=0) + ",")
            lines.append("};")
        print(name, len(co))
    with open(args.out, "w") as fh:
        fh.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
