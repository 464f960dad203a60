"""Fit and check of the branch-free erf used by the GELU epilogues (tc_common.cuh: gelu_fit).

erf(t) = 1 - 2^(-t*g(t)), g = weighted minimax (Lawson) polynomial of -log2(erfc(t))/t on [0,4]; prints the max abs error
of the fp32 evaluation for degrees 6..9, the degree-7 coefficients, and the error of the resulting fp32 GELU against
float64 next to that of the erff-based fp32 form.  CPU only (numpy + scipy)."""
import numpy as np
from scipy.special import erf, erfc
import numpy.polynomial.chebyshev as C
# g(t) = -log2(erfc(t)) / t  on [0, T];  erf(t) = 1 - 2^(-t*g(t))
T = 4.0
def g(t):
    t = np.asarray(t, dtype=np.float64)
    out = np.empty_like(t)
    small = t < 1e-8
    out[small] = 2/np.sqrt(np.pi)/np.log(2)
    ts = t[~small]
    out[~small] = -np.log2(erfc(ts))/ts
    return out
# weighted least squares on Chebyshev nodes; weight ~ sensitivity: d erf = erfc * ln2 * t * dg
N = 4000
x = np.cos(np.pi*(np.arange(N)+0.5)/N)
t = (x+1)*T/2
w = erfc(t)*np.log(2)*np.maximum(t,1e-3)
best=None
for deg in (6,7,8,9):
    V = np.vander(t, deg+1, increasing=True)
    coef, *_ = np.linalg.lstsq(V*w[:,None], g(t)*w, rcond=None)
    # iterate reweighting (Lawson) for minimax
    lw = np.ones_like(t)
    for it in range(60):
        coef, *_ = np.linalg.lstsq(V*(w*lw)[:,None], g(t)*w*lw, rcond=None)
        err = np.abs((V@coef - g(t))*w)
        lw = lw*(err/err.mean())**0.5
        lw /= lw.mean()
    # evaluate in float32 emulation
    tt = np.linspace(0, 6, 2000001).astype(np.float32)
    c32 = coef.astype(np.float32)
    tc = np.minimum(tt, np.float32(T))
    p = np.full_like(tc, c32[-1])
    for c in c32[-2::-1]:
        p = (p*tc + c).astype(np.float32)
    e = np.exp2((-(tc*p)).astype(np.float32).astype(np.float64)).astype(np.float32)
    ap = (np.float32(1)-e).astype(np.float32)
    ref = erf(tt.astype(np.float64))
    print(deg, 'max abs err', np.max(np.abs(ap-ref)), 'at', tt[np.argmax(np.abs(ap-ref))])
    if deg==8: best=coef
print(repr(best))

# degree 7 coefficients, GELU emulation in fp32
deg=7
V = np.vander(t, deg+1, increasing=True)
lw = np.ones_like(t)
for it in range(80):
    coef, *_ = np.linalg.lstsq(V*(w*lw)[:,None], g(t)*w*lw, rcond=None)
    err = np.abs((V@coef - g(t))*w)
    lw = lw*(err/err.mean())**0.5; lw/=lw.mean()
c32 = coef.astype(np.float32)
print('coef7 =', [float(c) for c in c32])
x = np.linspace(-9, 9, 4000001).astype(np.float32)
f32=np.float32
tt = (np.abs(x)*f32(0.70710678118654752440)).astype(f32)
tc = np.minimum(tt, f32(4.0))
p = np.full_like(tc, c32[-1])
for c in c32[-2::-1]:
    p = (p*tc + c).astype(f32)
z = (-(tc*p)).astype(f32)
e = np.exp2(z.astype(np.float64)).astype(f32)
h = (e*f32(-0.5)+f32(0.5)).astype(f32)
gl = (np.abs(x)*h + (f32(0.5)*x).astype(f32)).astype(f32)
ref = 0.5*x.astype(np.float64)*(1+erf(x.astype(np.float64)/np.sqrt(2)))
# exact-erf fp32 path emulation: erf rounded to fp32 then fp32 ops
er32 = erf((x*f32(0.70710678118654752440)).astype(f32).astype(np.float64)).astype(f32)
g32 = (x*(f32(0.5)*(f32(1.0)+er32)).astype(f32)).astype(f32)
print('fit  gelu max abs err', np.max(np.abs(gl-ref)), 'rel-to-|x|', np.max(np.abs(gl-ref)/np.maximum(np.abs(x),1e-3)))
print('erff gelu max abs err', np.max(np.abs(g32-ref)), 'rel-to-|x|', np.max(np.abs(g32-ref)/np.maximum(np.abs(x),1e-3)))
