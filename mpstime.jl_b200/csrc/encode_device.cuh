// Device-side basis functions (reference Encodings/bases.jl:13-92).  Real bases write d doubles,
// complex bases write (re, im) interleaved, 2*d doubles.
#pragma once
#include "mpst_common.cuh"

// Legendre, L2-normalised: phi_l(x) = sqrt((2l+1)/2) P_l(x), Bonnet recurrence
// (bases.jl:77-84 via LegendrePolynomials.Pl(...; norm=Val(:normalized))).
__device__ __forceinline__ void legendre_point(double x, int d, bool norm, double* v) {
    double pm = 1.0, p = x;
    v[0] = 0.70710678118654757;                        // sqrt(1/2)
    if (d > 1) v[1] = 1.2247448713915889 * x;          // sqrt(3/2) x
    for (int l = 1; l < d - 1; l++) {
        double pn = ((double)(2 * l + 1) * x * p - (double)l * pm) / (double)(l + 1);
        pm = p;
        p = pn;
        v[l + 1] = sqrt((double)(2 * l + 3) * 0.5) * p;
    }
    if (norm) {                                        // bases.jl:86-89
        double s = 1.0 / sqrt(sqrt((double)(2 * d + 1) * 0.5) * (double)d);
        for (int l = 0; l < d; l++) v[l] *= s;
    }
}

template <int BASIS>
__device__ __forceinline__ void encode_point(double x, int d, double* v) {
    if (BASIS == MPST_BASIS_LEGENDRE_NO_NORM) {
        legendre_point(x, d, false, v);
    } else if (BASIS == MPST_BASIS_LEGENDRE_NORM) {
        legendre_point(x, d, true, v);
    } else if (BASIS == MPST_BASIS_FOURIER) {          // bases.jl:23-42: cispi(f_k x)/sqrt(d)
        const double s = 1.0 / sqrt((double)d);
        for (int k = 0; k < d; k++) {
            int f = (k == 0) ? 0 : ((k & 1) ? (k + 1) / 2 : -(k / 2));
            double sn, cs;
            sincospi((double)f * x, &sn, &cs);
            v[2 * k] = cs * s;
            v[2 * k + 1] = sn * s;
        }
    } else if (BASIS == MPST_BASIS_STOUDENMIRE) {      // bases.jl:13-20
        double sn, cs, s2, c2;
        sincospi(1.5 * x, &sn, &cs);
        sincospi(0.5 * x, &s2, &c2);
        v[0] = cs * c2; v[1] = sn * c2;
        v[2] = cs * s2; v[3] = -sn * s2;
    } else if (BASIS == MPST_BASIS_SAHAND) {           // bases.jl:53-74
        const double dx = 2.0 / (double)d;
        for (int i = 1; i <= d; i++) {
            double interval = ceil((double)i / 2.0);
            double startx = (interval - 1.0) * dx;
            double re = 0.0, im = 0.0;
            if (startx <= x && x <= interval * dx) {
                double sn, cs, s2, c2;
                sincospi(1.5 * x / dx, &sn, &cs);
                sincospi(0.5 * (x - startx) / dx, &s2, &c2);
                if (i & 1) { re = cs * c2; im = sn * c2; }
                else       { re = cs * s2; im = -sn * s2; }
            }
            v[2 * (i - 1)] = re;
            v[2 * (i - 1) + 1] = im;
        }
    } else {                                           // uniform, bases.jl:2-5
        for (int k = 0; k < d; k++) v[k] = 1.0 / (double)d;
    }
}
