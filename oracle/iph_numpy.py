"""A SECOND, independent restatement of the Quemerais IPH model -- TEST INFRASTRUCTURE ONLY.

Written directly from the Fortran (reference src/quemerais_IPH_model/ipbackgroundCFR_fun.f: BACKGROUND :176-320,
T :366-397, TOP :399-490, DEN :492-548, INTENSM_PH :550-657, IPAL3M :659-741), by a different route than
oracle/iph_oracle.c: every line of sight is marched at once with vectorised numpy float32 arrays and activity masks, the
bracket searches are array comparisons instead of the Fortran's arithmetic-IF loops, and nothing is shared with the C
restatement or with the device kernel (csrc/iph.cu).  Its purpose is to catch transcription errors in iph_oracle.c:
tests/test_iph.py asserts that the two agree to 1e-5 on the reference's own table.

The Fortran itself cannot be compiled here (no gfortran in the image), so neither restatement has been run against it:
the IPH row stays "parity unpinned" -- what this file adds is that two independently written readings of the Fortran
agree.  Arithmetic is float32 throughout, in the Fortran's order of operations (REAL is 4 bytes; gfortran evaluates
REAL expressions in single precision on x86-64).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _f(x):
    return np.float32(x)


class IphNumpy:
    def __init__(self, tab, sin=np.sin, cos=np.cos, acos=np.arccos):
        """tab: the file's arrays as they stand (tests/golden/iph_real_table.npz; ipbackgroundCFR_fun.f:107-164).
        sin / cos / acos: the single-precision elementary functions to use.  They matter more than one would think: TOP
        ends its inner march on `SAB <= NORME` after ~20 steps of NORME/20, so a one-ulp change anywhere upstream (the
        wind-frame rotation, an angle) flips 20 steps into 21 and moves that step's optical depth by 5 % -- the model's
        output is only reproducible to ~1e-3 across math libraries.  The comparison with iph_oracle.c therefore runs
        with that file's elementary functions (glibc sinf / cosf, its documented acos polynomial); with numpy's own the
        two agree to the ~1e-3 the model allows."""
        self._acos = acos
        self.kmax, self.lmax = int(tab["kmax"]), int(tab["lmax"])
        self.ua = _f(1.4959E+11)                                           # :176
        self.dinf = np.asarray(tab["dinf_cm3"], dtype=F) * _f(1.E6)        # :177-179
        self.alt = np.asarray(tab["alt_au"], dtype=F) * self.ua            # :180-183 (ZALT == ALT)
        self.ang = np.asarray(tab["ang"], dtype=F)
        self.dans = np.asarray(tab["dans"], dtype=F)                       # [kmax][lmax]
        self.sot = np.asarray(tab["sot"], dtype=F)
        self.so = np.asarray(tab["so"], dtype=F)                           # [ninf][kmax][lmax]
        self.sn = np.asarray(tab["sn"], dtype=F)
        temp = _f(tab["temp"])
        # Lyman alpha constants, :190-221
        branch, xla, ptf = _f(1.), _f(1.21566E-05), _f(0.4162)
        am, bolk = _f(1.67333E-27), _f(1.38046E-23)
        py = _f(4) * np.arctan(_f(1.))
        self.dpi = py / _f(180.)
        spi = np.sqrt(py)
        xnuz = _f(1.) / xla
        e2, emas, c = _f(23.0677E-20), _f(9.1084E-28), _f(2.99793E+10)
        dldn = xla ** 2 * _f(1.E+08) / c
        sigman = py * e2 * ptf / (emas * c)
        self.sigmaf = sigman * dldn
        delnud = xnuz * np.sqrt(_f(2.) * bolk * temp / am) * _f(1.E+2)
        sig = sigman / (spi * delnud)
        self.sig = sig * _f(1.E-4)
        self.dtap = _f(1.) / self.sig
        # wind frame, :225-235
        alam, decv = _f(252.3) * self.dpi, _f(8.7) * self.dpi
        sa, ca, sd, cd = _f(sin(alam)), _f(cos(alam)), _f(sin(decv)), _f(cos(decv))
        self.A = np.array([[sa, -ca, _f(0.)], [cd * ca, cd * sa, sd], [-sd * ca, -sd * sa, cd]], dtype=F)

    # ---- bracket of an angle / a radius in the node arrays: what the arithmetic-IF loops of DEN / IPAL3M find
    def _angle_bracket(self, t):
        """-> (ll, llp, dt) 0-based; first J with T <= ANG(J): equality -> (J, J+1, 0), else (J-1, J, ratio)"""
        j = np.searchsorted(self.ang, t, side="left")                     # first index with ang[j] >= t
        j = np.minimum(j, self.lmax - 1)
        eq = self.ang[j] == t
        ll = np.where(eq, j, j - 1)
        llp = np.where(eq, np.minimum(j + 1, self.lmax - 1), j)            # (J+1 is multiplied by DT = 0 on equality)
        ll = np.maximum(ll, 0)
        den = np.where(eq, _f(1.), self.ang[j] - self.ang[np.maximum(j - 1, 0)])
        dt = np.where(eq, _f(0.), (t - self.ang[np.maximum(j - 1, 0)]) / den).astype(F)
        return ll, llp, dt

    def _radius_bracket(self, z):
        m = np.searchsorted(self.alt, z, side="left")
        m = np.minimum(m, self.kmax - 1)
        eq = self.alt[m] == z
        kk = np.where(eq, m, m - 1)
        kkp = np.where(eq, np.minimum(m + 1, self.kmax - 1), m)            # (M+1 is multiplied by DU = 0 on equality)
        kk = np.maximum(kk, 0)
        den = np.where(eq, _f(1.), self.alt[m] - self.alt[np.maximum(m - 1, 0)])
        du = np.where(eq, _f(0.), (z - self.alt[np.maximum(m - 1, 0)]) / den).astype(F)
        return kk, kkp, du

    @staticmethod
    def _bilinear(tab, kk, kkp, ll, llp, du, dt):
        fl = tab[kk, ll] + du * (tab[kkp, ll] - tab[kk, ll])
        flp = tab[kk, llp] + du * (tab[kkp, llp] - tab[kk, llp])
        return (fl + dt * (flp - fl)).astype(F)

    def den(self, z, t):
        """DEN (:492-548) -> (dan, ko 1-based, z as the caller sees it afterwards: clamped to ALT(KMAX))"""
        z = z.astype(F).copy()
        below = z < self.alt[0]
        z = np.where(~below & (z > self.alt[self.kmax - 1]), self.alt[self.kmax - 1], z)
        zz = np.where(below, self.alt[0], z)
        ll, llp, dt = self._angle_bracket(t)
        kk, kkp, du = self._radius_bracket(zz)
        dan = self._bilinear(self.dans, kk, kkp, ll, llp, du, dt)
        dan = np.where(below, _f(0.), dan)
        ko = np.where(below, 1, kk + 1)
        return dan.astype(F), ko, z

    def t_fun(self, to):
        """T(tau) (:366-397)"""
        to = to.astype(F)
        out = np.zeros_like(to)
        # tau > 2: trapezoid in x = 0.4 k
        depi = _f(2.) / np.sqrt(_f(3.14159265358))
        dx = _f(0.4)
        big = np.where(to < _f(600.), depi * np.exp(-np.minimum(to, _f(600.))) * _f(0.5) * dx, _f(0.)).astype(F)
        for k in range(1, 11):
            x = _f(k) * dx
            u = np.exp(-(x * x))
            uu = to * u
            d = np.where(uu < _f(600.), depi * u * np.exp(-np.minimum(uu, _f(600.))), _f(0.)).astype(F)
            big = (big + d * dx).astype(F)
        # tau <= 2: the series
        tn = np.ones_like(to)
        dtn = np.ones_like(to)
        q = _f(1.)
        while q < _f(12.):
            dtn = (-dtn * to / np.sqrt(q * (q + _f(1.)))).astype(F)
            tn = (tn + dtn).astype(F)
            q = q + _f(1.)
        out = np.where(to <= _f(2.), tn, big)
        return np.where(to < _f(0.), _f(0.), out).astype(F)

    def top(self, xf, yf, zf, xh, yh, zh, imd):
        """TOP (:399-490) for arrays of point pairs; imd 1-based density index"""
        ua = self.ua
        xa, xb, ya, yb, za, zb = xf / ua, xh / ua, yf / ua, yh / ua, zf / ua, zh / ua
        altp = self.alt[0] / ua
        ra = np.sqrt(xa * xa + ya * ya + za * za)
        rb = np.sqrt(xb * xb + yb * yb + zb * zb)
        zero = (ra <= altp) & (rb <= altp)
        swap = ra > rb
        xa, xb = np.where(swap, xb, xa), np.where(swap, xa, xb)
        ya, yb = np.where(swap, yb, ya), np.where(swap, ya, yb)
        za, zb = np.where(swap, zb, za), np.where(swap, za, zb)
        ra, rb = np.where(swap, rb, ra), np.where(swap, ra, rb)
        xab, yab, zab = xb - xa, yb - ya, zb - za
        norme = np.sqrt(xab * xab + yab * yab + zab * zab)
        zero |= norme < _f(.01)
        safe = np.where(zero, _f(1.), norme)
        xab, yab, zab = xab / safe, yab / safe, zab / safe
        dsa0 = norme / _f(20.)
        dinf = self.dinf[imd - 1]

        def step_limit(kp):
            kp = np.where(kp == self.kmax, kp - 1, kp)
            return (self.alt[kp] - self.alt[kp - 1]) / _f(3.) / ua            # (ALT(KP+1)-ALT(KP))/3./UA, 1-based KP

        with np.errstate(invalid="ignore", divide="ignore"):
            ta = self._acos(ya / ra) / self.dpi
        dn1, kp, _ = self.den(ra * ua, ta)
        dn1 = dinf * dn1
        dsab = np.minimum(step_limit(kp), dsa0)
        sab = np.zeros_like(ra)
        dt = np.zeros_like(ra)
        active = ~zero
        while active.any():
            xa = np.where(active, xa + dsab * xab, xa)
            ya = np.where(active, ya + dsab * yab, ya)
            za = np.where(active, za + dsab * zab, za)
            sab = np.where(active, sab + dsab, sab)
            ra = np.sqrt(xa * xa + ya * ya + za * za)
            with np.errstate(invalid="ignore", divide="ignore"):
                ta = self._acos(ya / ra) / self.dpi
            dn, kp, _ = self.den(ra * ua, ta)
            dn = dinf * dn
            dt = np.where(active, dt + (dn + dn1) * _f(.5) * dsab * self.sig * ua, dt).astype(F)
            dn1 = np.where(active, dn, dn1)
            dsab = np.where(active, np.minimum(step_limit(kp), dsa0), dsab)
            active = active & (sab <= norme)
        return np.where(zero, _f(0.), dt).astype(F)

    def ipal3m(self, r, t):
        """IPAL3M (:659-741) -> CT, FOO, F each [5][n]"""
        n = len(r)
        outside = (r < self.alt[0]) | (r >= self.alt[self.kmax - 1])
        rz = np.where(outside, self.alt[0], r)
        ll, llp, dt = self._angle_bracket(t)
        kk, kkp, du = self._radius_bracket(rz)
        ct, foo, f = np.zeros((5, n), F), np.zeros((5, n), F), np.zeros((5, n), F)
        cot = self._bilinear(self.sot, kk, kkp, ll, llp, du, dt)
        for i in range(len(self.dinf)):
            f[i] = np.where(outside, _f(0.), self._bilinear(self.sn[i], kk, kkp, ll, llp, du, dt))
            foo[i] = self._bilinear(self.so[i], kk, kkp, ll, llp, du, dt)      # (not zeroed on return in the Fortran)
            ct[i] = np.where(outside, _f(0.), cot * self.dinf[i])
        return ct, foo, f, outside

    def background(self, fs, pos, u1, v1, w1, want_steps=False):
        """BACKGROUND (:1-320) -> fln = xsn(2) in rayleigh [n]"""
        fs = _f(fs)
        gral = (fs * self.sigmaf) * _f(1.E-10)
        A, ua = self.A, self.ua
        x1, y1, z1 = (_f(p) for p in pos)
        x = (A[0, 0] * x1 + A[0, 1] * y1) * ua
        y = (A[1, 0] * x1 + A[1, 1] * y1 + A[1, 2] * z1) * ua
        z = (A[2, 0] * x1 + A[2, 1] * y1 + A[2, 2] * z1) * ua
        u1, v1, w1 = (np.asarray(a, dtype=F) for a in (u1, v1, w1))
        u = A[0, 0] * u1 + A[0, 1] * v1
        v = A[1, 0] * u1 + A[1, 1] * v1 + A[1, 2] * w1
        w = A[2, 0] * u1 + A[2, 1] * v1 + A[2, 2] * w1
        n = len(u)
        idb = 1
        tt = np.zeros((5, n), F)
        fln = np.zeros((5, n), F)
        s = np.zeros(n, F)
        steps = np.zeros(n, np.int32)
        rr = np.sqrt(x * x + y * y + z * z)
        active = np.full(n, bool(rr <= self.alt[self.kmax - 1]))
        yp = np.full(n, y, F)
        r = np.full(n, rr, F)
        xav, yav, zav = np.full(n, x, F), np.full(n, y, F), np.full(n, z, F)
        foo_keep = np.zeros((5, n), F)
        while active.any():
            with np.errstate(invalid="ignore", divide="ignore"):
                teta = self._acos(yp / r) / self.dpi
            dna, ko, r = self.den(r, teta)                       # DEN clamps R in the caller's variable
            dn1 = self.dinf[idb - 1] * dna
            dn1 = np.where(dn1 == _f(0.), _f(1.), dn1)
            dp = self.dtap * _f(0.05) / dn1
            ko_c = np.where(ko < self.kmax, ko, self.kmax - 1)
            dua = (self.alt[ko_c] - self.alt[ko_c - 1]) / _f(2.)   # (ALT(KO+1)-ALT(KO))/2, or the last interval
            dp = np.maximum(np.minimum(dp, dua), ua / _f(10.)).astype(F)
            s = np.where(active, s + dp, s).astype(F)
            xp, ypn, zp = x + s * u, y + s * v, z + s * w
            yp = np.where(active, ypn, yp).astype(F)
            rn = np.sqrt(xp * xp + ypn * ypn + zp * zp)
            r = np.where(active, rn, r).astype(F)
            steps += active
            active = active & ~(r > self.alt[self.kmax - 1])
            if not active.any():
                break
            with np.errstate(invalid="ignore", divide="ignore"):
                teta = self._acos(yp / r) / self.dpi
            ct, foo, f, outside = self.ipal3m(r, teta)
            foo_keep = np.where(outside[None, :], foo_keep, foo)   # FOO keeps its previous value when IPAL3M returns early
            dtt = self.top(xav, yav, zav, xp.astype(F), ypn.astype(F), zp.astype(F), idb)
            cosff = (u * xp + v * ypn + w * zp) / r
            corec = _f(0.25) * cosff * cosff + (_f(11.) / _f(12.))
            for ii in range(5):
                tt[ii] = np.where(active, tt[ii] + dtt * self.dinf[ii] / self.dinf[idb - 1], tt[ii]).astype(F)
                ffnn = f[ii] + foo_keep[ii] * (corec - _f(1.))
                dflnc = ffnn * gral * self.t_fun(tt[ii]) * dp
                fln[ii] = np.where(active, fln[ii] + dflnc, fln[ii]).astype(F)
            xav = np.where(active, xp, xav).astype(F)
            yav = np.where(active, ypn, yav).astype(F)
            zav = np.where(active, zp, zav).astype(F)
        return (fln[1], steps) if want_steps else fln[1]
