/*
 * bs_oracle_impl.h -- TEST INFRASTRUCTURE ONLY.  Included twice by bs_oracle.c, once per precision.
 *
 * Before including, define:
 *   FP        the reference's `fptype` (float or double)              blackscholes.c:85
 *   SFX(name) token-pasting macro that appends _f32 / _f64
 *   FP_EXP / FP_LOG / FP_SQRT / FP_ABS   the libm entry the reference's C++ overload set resolves to
 *             for an `fptype` argument (expf/logf/sqrtf/fabsf when fptype=float: verified by
 *             disassembly of the reference build)
 *   FP_SCAN   the fscanf conversion ("%f" or "%lf")
 *
 * PRECISION RULE.  The reference writes its constants as *double* literals.  With fptype=float every
 * expression that touches such a literal is therefore evaluated in double and rounded back to float
 * on assignment, while literal-free expressions stay in float.  This file keeps every literal's type
 * and every assignment boundary of blackscholes.c:128-258 so that plain C arithmetic conversion
 * reproduces those roundings bit for bit (FLT_EVAL_METHOD==0 on x86-64, contraction off).
 */

/* Normal CDF, Abramowitz-Stegun 26.2.17 as used by Hull -- follows blackscholes.c:128-184. */
FP SFX(bs_oracle_cndf)(FP x)
{
    const int negative = (x < 0.0);                 /* :144-148  compare against a double literal   */
    FP ax, npx, k1, k2, k3, k4, k5, acc, term, lead, out;

    if (negative) x = -x;
    ax = x;

    npx = FP_EXP(-0.5f * x * x);                    /* :152  float literal: stays in fptype          */
    npx = npx * 0.39894228040143270286;             /* :154  double literal (inv_sqrt_2xPI, :126)    */

    k1 = 0.2316419 * ax;                            /* :156 */
    k1 = 1.0 + k1;                                  /* :157 */
    k1 = 1.0 / k1;                                  /* :158 */
    k2 = k1 * k1;                                   /* :159-162  successive powers in fptype         */
    k3 = k2 * k1;
    k4 = k3 * k1;
    k5 = k4 * k1;

    lead = k1 * 0.319381530;                        /* :164 */
    acc  = k2 * (-0.356563782);                     /* :165 */
    term = k3 * 1.781477937;                        /* :166 */
    acc  = acc + term;                              /* :167 */
    term = k4 * (-1.821255978);                     /* :168 */
    acc  = acc + term;                              /* :169 */
    term = k5 * 1.330274429;                        /* :170 */
    acc  = acc + term;                              /* :171 */

    lead = acc + lead;                              /* :173 */
    out  = lead * npx;                              /* :174 */
    out  = 1.0 - out;                               /* :175 */

    if (negative) out = 1.0 - out;                  /* :179-181 */
    return out;
}

/* European option, no dividends -- follows blackscholes.c:190-258 (`timet` is unused there). */
FP SFX(bs_oracle_price)(FP spot, FP strike, FP rate, FP vol, FP tte, int otype)
{
    FP sqrt_t, log_sk, pw, d1, d2, den, fv, n1, n2, m1, m2, price;

    sqrt_t = FP_SQRT(tte);                          /* :224 */
    log_sk = FP_LOG(spot / strike);                 /* :226 */

    pw = vol * vol;                                 /* :231 */
    pw = pw * 0.5;                                  /* :232  double literal                          */

    d1 = rate + pw;                                 /* :234 */
    d1 = d1 * tte;                                  /* :235 */
    d1 = d1 + log_sk;                               /* :236 */

    den = vol * sqrt_t;                             /* :238 */
    d1 = d1 / den;                                  /* :239 */
    d2 = d1 - den;                                  /* :240 */

    n1 = SFX(bs_oracle_cndf)(d1);                   /* :245 */
    n2 = SFX(bs_oracle_cndf)(d2);                   /* :246 */

    fv = strike * (FP_EXP(-(rate) * (tte)));        /* :248 */
    if (otype == 0) {                               /* :249  0 = call                                */
        price = (spot * n1) - (fv * n2);            /* :250 */
    } else {
        m1 = (1.0 - n1);                            /* :252 */
        m2 = (1.0 - n2);                            /* :253 */
        price = (fv * m2) - (spot * m1);            /* :254 */
    }
    return price;
}

/* The Map body over the SoA streams -- blackscholes.c:328-331 (FF), :634-637 (bs_thread).
 * One pass; callers repeat it NUM_RUNS times if they want the reference's ROI cost. */
void SFX(bs_oracle_map)(size_t n, const FP *spot, const FP *strike, const FP *rate, const FP *vol,
                        const FP *tte, const int *otype, FP *prices, int nthreads)
{
    long i;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (i = 0; i < (long)n; i++)
        prices[i] = SFX(bs_oracle_price)(spot[i], strike[i], rate[i], vol[i], tte[i], otype[i]);
}

/* ERR_CHK -- blackscholes.c:333-340.  Counts |DGrefval - price| >= 1e-4 for ONE pass and records
 * the first `cap` offending indices.  (The reference bumps numError once per run, so its final
 * "Num Errors" line is NUM_RUNS times this count.) */
unsigned long long SFX(bs_oracle_errchk)(size_t n, const FP *prices, const FP *refval,
                                         long long *bad_idx, size_t cap)
{
    unsigned long long bad = 0;
    size_t i;
    for (i = 0; i < n; i++) {
        FP delta = refval[i] - prices[i];
        if (FP_ABS(delta) >= 1e-4) {
            if (bad_idx && bad < cap) bad_idx[bad] = (long long)i;
            bad++;
        }
    }
    return bad;
}

/* Loader -- blackscholes.c:696-739 (header "%i", rows of nine conversions) fused with the
 * AoS->SoA staging of :760-767 (only 'P' is a put).  Returns the number of options, or -1 when the
 * file cannot be opened / -2 when a conversion fails (the reference prints and exits there).
 * When the output pointers are NULL only the header is read. */
long SFX(bs_oracle_load)(const char *path, size_t cap, FP *spot, FP *strike, FP *rate, FP *divq,
                         FP *vol, FP *tte, int *otype, FP *divs, FP *refval)
{
    FILE *f = fopen(path, "r");
    int n = 0, rv;
    long i;
    if (!f) return -1;
    rv = fscanf(f, "%i", &n);
    if (rv != 1) { fclose(f); return -2; }
    if (!spot) { fclose(f); return n; }
    if ((size_t)n > cap) n = (int)cap;
    for (i = 0; i < n; i++) {
        FP a, b, c, d, e, g, h, j;
        char ty;
        rv = fscanf(f, FP_SCAN " " FP_SCAN " " FP_SCAN " " FP_SCAN " " FP_SCAN " " FP_SCAN " %c " FP_SCAN " " FP_SCAN,
                    &a, &b, &c, &d, &e, &g, &ty, &h, &j);
        if (rv != 9) { fclose(f); return -2; }
        spot[i] = a; strike[i] = b; rate[i] = c; vol[i] = e; tte[i] = g;
        if (divq) divq[i] = d;
        if (divs) divs[i] = h;
        if (refval) refval[i] = j;
        otype[i] = (ty == 'P') ? 1 : 0;
    }
    fclose(f);
    return n;
}

/* Writer -- blackscholes.c:923-947: "%i\n" then one "%.18f\n" per price. */
int SFX(bs_oracle_write)(const char *path, size_t n, const FP *prices)
{
    FILE *f = fopen(path, "w");
    size_t i;
    if (!f) return -1;
    if (fprintf(f, "%i\n", (int)n) < 0) { fclose(f); return -2; }
    for (i = 0; i < n; i++)
        if (fprintf(f, "%.18f\n", prices[i]) < 0) { fclose(f); return -2; }
    return fclose(f) == 0 ? 0 : -3;
}
