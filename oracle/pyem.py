"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's length-based EM for the stutter model
(EMStutterGenotyper, /root/reference/src/em_stutter_genotyper.{h,cpp}; StutterModel, src/stutter_model.{h,cpp}; the posterior
pass of src/genotyper.cpp:45-83; mathops.cpp; fastonebigheader.h:188-218, 320-357).  Pure-Python loops in double precision with
numpy float32 for the single-precision approximations: small cases only.  Pinned by the reference itself compiled in place
(oracle/em_driver.cpp -> oracle/_ref/libltr_ref_em.so; tests/test_oracle_em.py) and by tests/golden/em.json recorded from it.
Checker of the CUDA kernel longtr_b200/csrc/em_kernel.cu (tests/test_gpu_em.py)."""
import math

import numpy as np

F = np.float32
LOG_ONE_HALF = math.log(0.5)      # mathops.cpp:10
TOLERANCE = 1e-10                 # mathops.cpp:11
LOG_THRESH = math.log(0.001)      # mathops.h:36
DBL_MAX = 1.7976931348623157e308


def int_log(v):                   # mathops.cpp:14-22
    return -1000.0 if v == 0 else math.log(v)


def _u32(f):
    return int(np.array(f, dtype=np.float32).view(np.uint32))


def _f32(u):
    return F(np.array(u & 0xFFFFFFFF, dtype=np.uint32).view(np.float32))


def fastpow2(p):                  # fastonebigheader.h:188-199
    p = F(p)
    offset = F(1.0) if p < 0 else F(0.0)
    clipp = F(-126.0) if p < -126 else p
    w = int(clipp)                # truncation
    z = F(F(clipp - F(w)) + offset)
    t = F(F(F(clipp + F(121.2740575)) + F(F(27.7280233) / F(F(4.84252568) - z))) - F(F(1.49012907) * z))
    return _f32(int(F(F(8388608.0) * t)))


def fastexp(p):                   # :201-205
    return fastpow2(F(F(1.442695040) * F(p)))


def fastlog(x):                   # :320-337
    bits = _u32(F(x))
    mx = _f32((bits & 0x007FFFFF) | 0x3f000000)
    y = F(F(bits) * F(1.1920928955078125e-7))
    l2 = F(F(F(y - F(124.22551499)) - F(F(1.498030302) * mx)) - F(F(1.72587999) / F(F(0.3520887068) + mx)))
    return F(F(0.69314718) * l2)


def fasterexp(p):                 # :207-218
    x = F(F(1.442695040) * F(p))
    clipp = F(-126.0) if x < -126 else x
    return _f32(int(F(F(8388608.0) * F(clipp + F(126.94269504)))))


def fasterlog(x):                 # :347-357
    y = F(F(_u32(F(x))) * F(8.2629582881927490e-8))
    return F(y - F(87.989971088))


def log_sum_exp(vals):            # mathops.cpp:45-51, 65-71
    m = max(vals)
    total = 0.0
    for v in vals:
        total += math.exp(v - m)
    return m + math.log(total)


def log_sum_exp2(a, b):           # :53-58
    return a + math.log(1 + math.exp(b - a)) if a > b else b + math.log(1 + math.exp(a - b))


def log_sum_exp3(a, b, c):        # :60-63
    m = max(max(a, b), c)
    return m + math.log(math.exp(a - m) + math.exp(b - m) + math.exp(c - m))


def fast_log_sum_exp2(a, b):      # :87-96
    if a > b:
        d = b - a
        return a if d < LOG_THRESH else a + float(fastlog(F(1) + fastexp(d)))
    d = a - b
    return b if d < LOG_THRESH else b + float(fastlog(F(1) + fastexp(d)))


def fast_log_sum_exp(vals):       # :98-107
    m = max(vals)
    total = 0.0
    for v in vals:
        d = v - m
        if d > LOG_THRESH:
            total += float(fasterexp(d))
    return m + float(fasterlog(total))


class StutterModel:               # stutter_model.h:17-67, stutter_model.cpp:29-53
    def __init__(self, ig, iu, idn, og, ou, odn, motif_len):
        self.p = (ig, iu, idn, og, ou, odn)
        self.in_log_step, self.in_log_nostep = math.log(1 - ig), math.log(ig)
        self.in_log_up, self.in_log_down = math.log(iu), math.log(idn)
        self.out_log_step, self.out_log_nostep = math.log(1 - og), math.log(og)
        self.out_log_up, self.out_log_down = math.log(ou), math.log(odn)
        self.log_equal = math.log(1 - iu - idn - ou - odn)
        self.motif_len = motif_len

    def log_stutter_pmf(self, sample_bps, read_bps):
        bp_diff = read_bps - sample_bps
        q = int(bp_diff / self.motif_len)        # C division truncates toward zero
        if bp_diff - q * self.motif_len != 0:
            eff = bp_diff - q
            if eff < 0:
                return self.out_log_down + self.out_log_nostep + self.out_log_step * (-eff - 1)
            return self.out_log_up + self.out_log_nostep + self.out_log_step * (eff - 1)
        if q == 0:
            return self.log_equal
        if q < 0:
            return self.in_log_down + self.in_log_nostep + self.in_log_step * (-q - 1)
        return self.in_log_up + self.in_log_nostep + self.in_log_step * (q - 1)

    def within(self, o, md):
        return all(abs(a - b) < md for a, b in zip(self.p, o.p))


def em_train(reads_per_sample, bp_diff, log_p1, log_p2, motif_len, haploid=False, max_iter=100, abs_conv=0.01,
             frac_conv=0.001):
    """EMStutterGenotyper(haploid, motif, num_bps, log_p1, log_p2, names, ref_allele = 0).train(max_iter, abs_conv, frac_conv).
    Reads sample-major.  -> dict(trained, params, n_iter, lls, log_gt_priors, alleles)."""
    S = len(reads_per_sample)
    sample = [s for s in range(S) for _ in range(reads_per_sample[s])]
    R = len(sample)
    sizes = sorted(set(int(b) for b in bp_diff) - {0})           # em_stutter_genotyper.h:61-77
    bps = [0] + sizes
    A = len(bps)
    aidx = [bps.index(int(b)) for b in bp_diff]
    # init_log_gt_priors (:10-19)
    pri = [1.0] * A
    for r in range(R):
        pri[aidx[r]] += 1.0 / reads_per_sample[sample[r]]
    tot = 0.0
    for x in pri:
        tot += x
    log_total = math.log(tot)
    pri = [math.log(x) - log_total for x in pri]
    model = StutterModel(0.9, 0.1, 0.1, 0.8, 0.01, 0.01, motif_len)   # :57-60
    LL, num_iter, lls = -DBL_MAX, 1, []
    trained = False
    while num_iter <= max_iter:
        # E step: calc_hap_aln_probs (:140-144)
        aln = [[model.log_stutter_pmf(bps[a], bps[aidx[r]]) for a in range(A)] for r in range(R)]
        # calc_log_sample_posteriors (genotyper.cpp:45-83) with init_log_sample_priors (:128-138)
        post = [[[(pri[i] + pri[j]) if not haploid else (pri[i] if i == j else -DBL_MAX / 2) for j in range(A)]
                 for i in range(A)] for _ in range(S)]
        for r in range(R):
            ps = post[sample[r]]
            for i in range(A):
                for j in range(A):
                    if aln[r][i] < -600:
                        aln[r][i] = -600
                    if aln[r][j] < -600:
                        aln[r][j] = -600
                    ps[i][j] += math.log(math.exp(aln[r][i] + log_p1[r] + LOG_ONE_HALF) +
                                         math.exp(aln[r][j] + log_p2[r] + LOG_ONE_HALF))
        new_LL = 0.0
        for s in range(S):
            flat = [post[s][i][j] for i in range(A) for j in range(A)]
            t = log_sum_exp(flat)
            for i in range(A):
                for j in range(A):
                    post[s][i][j] -= t
            new_LL += t
        # recalc_log_read_phase_posteriors (:146-163)
        phase = [[[None] * A for _ in range(A)] for _ in range(R)]
        for r in range(R):
            for i in range(A):
                for j in range(A):
                    one = LOG_ONE_HALF + log_p1[r] + model.log_stutter_pmf(bps[i], bps[aidx[r]])
                    two = LOG_ONE_HALF + log_p2[r] + model.log_stutter_pmf(bps[j], bps[aidx[r]])
                    t = fast_log_sum_exp2(one, two)
                    phase[r][i][j] = (one - t, two - t)
        lls.append(new_LL)
        if new_LL < LL + TOLERANCE:                                   # :196-200
            trained = True
            break
        # M step: recalc_log_gt_priors (:21-55)
        mx, tl = [-DBL_MAX / 2] * A, [0.0] * A

        def upd(v, k):
            if v <= mx[k]:
                tl[k] += math.exp(v - mx[k])
            else:
                tl[k] *= math.exp(mx[k] - v)
                tl[k] += 1.0
                mx[k] = v
        for s in range(S):
            for i in range(A):
                upd(log_sum_exp(post[s][i]), i)
        for s in range(S):
            for i in range(A):
                for j in range(A):
                    upd(post[s][i][j], j)
        pri = [mx[k] + math.log(tl[k]) for k in range(A)]
        log_total = log_sum_exp(pri)
        pri = [x - log_total for x in pri]
        # recalc_stutter_model (:62-126)
        in_up, in_down, in_eq, in_diffs = [0.0], [0.0], [0.0], [0.0, math.log(1.1)]
        out_up, out_down, out_diffs = [0.0], [0.0], [0.0, math.log(1.1)]
        for r in range(R):
            for i in range(A):
                for j in range(A):
                    for ph in range(2):
                        gt = i if ph == 0 else j
                        d = bps[aidx[r]] - bps[gt]
                        factor = post[sample[r]][i][j] + phase[r][i][j][ph]
                        if d == 0:
                            in_eq.append(factor)
                            continue
                        q = int(d / motif_len)
                        if d - q * motif_len != 0:
                            out_diffs.append(factor + int_log(abs(d - q)))
                            (out_up if d > 0 else out_down).append(factor)
                        else:
                            in_diffs.append(factor + int_log(abs(q)))
                            (in_up if d > 0 else in_down).append(factor)
        t_in_up, t_in_down, t_in_eq = fast_log_sum_exp(in_up), fast_log_sum_exp(in_down), fast_log_sum_exp(in_eq)
        t_in_diffs = fast_log_sum_exp(in_diffs)
        t_out_up, t_out_down, t_out_diffs = fast_log_sum_exp(out_up), fast_log_sum_exp(out_down), fast_log_sum_exp(out_diffs)
        out_total = fast_log_sum_exp2(t_out_up, t_out_down)
        in_pgeom = min(0.999, math.exp(log_sum_exp2(t_in_up, t_in_down) - t_in_diffs))
        out_pgeom = min(0.999, math.exp(out_total - t_out_diffs))
        log_total = log_sum_exp2(log_sum_exp3(t_in_up, t_in_down, t_in_eq), out_total)
        prev = model
        model = StutterModel(in_pgeom, math.exp(t_in_up - log_total), math.exp(t_in_down - log_total), out_pgeom,
                             math.exp(t_out_up - log_total), math.exp(t_out_down - log_total), motif_len)
        abs_change = new_LL - LL
        frac_change = -(new_LL - LL) / LL
        if (abs_change < abs_conv and frac_change < frac_conv) or model.within(prev, 0.0001):   # :210-222
            trained = True
            break
        LL = new_LL
        num_iter += 1
    return dict(trained=trained, params=np.array(model.p), n_iter=len(lls), lls=np.array(lls), log_gt_priors=np.array(pri),
                alleles=bps)
