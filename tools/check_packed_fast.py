"""Exhaustive CPU check of the packed 16x2 compare identities proposed for k_fast's stage-1 walk:
two pixels per 32-bit register, bytes widened to 0x0100 | p by PRMT, per-half predicates read from bit 15 without
cross-half borrows.
  bright half:  (max2(max2(dn, up)... ) folded to one value m)   m > v + t   <=>  bit15( m' + (0x8000 - (v' + t + 1)) )
  dark half:    n < v - t                                          <=>  bit15( (v' - t - 1 + 0x8000) - n' )
with x' = x | 0x100 (x in 0..255).  All (m, v, t) are enumerated; halves are checked jointly on random pairs to prove that no
borrow / carry crosses bit 16."""
import numpy as np

m = np.arange(256, dtype=np.int64)[:, None, None]
v = np.arange(256, dtype=np.int64)[None, :, None]
t = np.arange(256, dtype=np.int64)[None, None, :]
mp, vp = m | 0x100, v | 0x100
e = (mp + (0x8000 - (vp + t + 1))) & 0xffff
assert (((e >> 15) & 1) == (m > v + t)).all()
f = ((vp - t - 1 + 0x8000) - mp) & 0xffff
assert (((f >> 15) & 1) == (m < v - t)).all()
# every intermediate stays inside 16 bits without wrapping, so two halves can share one 32-bit add / sub
assert (mp + (0x8000 - (vp + t + 1)) >= 0).all() and (mp + (0x8000 - (vp + t + 1)) < 0x10000).all()
assert ((vp - t - 1 + 0x8000) - mp >= 0).all() and ((vp - t - 1 + 0x8000) - mp < 0x10000).all()
rng = np.random.default_rng(0)
a = rng.integers(0, 256, (4, 1_000_000)); tt = rng.integers(0, 256, 1_000_000)
m2 = (a[0] | 0x100) | ((a[1] | 0x100) << 16); v2 = (a[2] | 0x100) | ((a[3] | 0x100) << 16)
k1 = (0x80008000 - (tt + 1) * 0x10001) - v2
e2 = (m2 + k1) & 0xffffffff
assert ((((e2 >> 15) & 1) == (a[0] > a[2] + tt)) & (((e2 >> 31) & 1) == (a[1] > a[3] + tt))).all()
k2 = v2 + (0x80008000 - (tt + 1) * 0x10001)
f2 = (k2 - m2) & 0xffffffff
assert ((((f2 >> 15) & 1) == (a[0] < a[2] - tt)) & (((f2 >> 31) & 1) == (a[1] < a[3] - tt))).all()
print("packed 16x2 compare identities hold for all (m, v, t) and for packed pairs")
