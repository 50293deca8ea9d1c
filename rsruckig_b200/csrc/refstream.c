// refstream.c — the input stream of the reference's own benchmark (bench/src/benchmark_target.rs:29-70, :115-153), so that
// the same 7-DoF problems can be fed to rsruckig (tools/rsruckig_dump), to the oracle and to the GPU. Host-only C, built by
// __graft_entry__.build() into librefstream.so; a workload generator, not part of the calculation path.
//
// The reference draws from three `Pcg64Mcg::seed_from_u64(seed)` streams (seeds 42 / 43 / 44: positions, dynamics, limits)
// through `rand_distr::Normal` / `rand::distributions::Uniform`. Those crates are not vendored in /root/reference
// (Cargo.lock: rand 0.8.5, rand_core 0.6.4, rand_pcg 0.3.1, rand_distr 0.4.3), so their published algorithms are restated:
//   * rand_core 0.6 `SeedableRng::seed_from_u64`: a PCG32 sequence (multiplier 6364136223846793005, increment
//     11634580027462260723) fills the 16 seed bytes, four little-endian bytes at a time;
//   * rand_pcg 0.3 `Mcg128Xsl64` (= Pcg64Mcg): state = seed | 1, state *= 0x2360ED051FC65DA44385DF649FCCF645 per draw,
//     output XSL-RR (xor of the two halves rotated right by the top six bits);
//   * rand 0.8 `Uniform<f64>`: (u64 >> 12) as the mantissa of a double in [1, 2), minus 1, times `scale` plus `low`, with
//     `scale` decreased by ulps until the largest value stays below `high`;
//   * rand_distr 0.4 `StandardNormal`: the 256-layer ziggurat (ZIGNOR, Doornik 2005) on `next_u64` — low 8 bits pick the layer,
//     the top 52 bits make u in [-1, 1) — with the tail drawn by two Open01 logarithms, and the rejection test on
//     `Standard` f64 (53 bits); the layer tables are those of rand's `ziggurat_tables.py` (R = 3.6541528853610088,
//     V = 0.00492867323399), printed with 18 decimals and parsed back exactly as the crate's literals are.
// Known answers (rand_pcg's own tests): see tests/test_refstream.py.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

typedef struct { u128 state; } pcg64mcg;

static const uint64_t MULT_HI = 0x2360ED051FC65DA4ULL, MULT_LO = 0x4385DF649FCCF645ULL;

static void pcg_new(pcg64mcg* r, u128 state) { r->state = state | 1; }

static uint64_t pcg_next_u64(pcg64mcg* r) {
    const u128 mult = ((u128)MULT_HI << 64) | MULT_LO;
    r->state *= mult;
    const u128 s = r->state;
    const uint32_t rot = (uint32_t)(s >> 122);
    const uint64_t xsl = (uint64_t)(s >> 64) ^ (uint64_t)s;
    return (xsl >> rot) | (xsl << ((64 - rot) & 63));
}

static void pcg_from_seed(pcg64mcg* r, const uint8_t seed[16]) {
    uint64_t w[2];
    for (int k = 0; k < 2; ++k) {
        w[k] = 0;
        for (int b = 0; b < 8; ++b) w[k] |= (uint64_t)seed[8 * k + b] << (8 * b);
    }
    pcg_new(r, (u128)w[0] | ((u128)w[1] << 64));
}

static void pcg_seed_from_u64(pcg64mcg* r, uint64_t state) {
    uint8_t seed[16];
    for (int c = 0; c < 4; ++c) {
        state = state * 6364136223846793005ULL + 11634580027462260723ULL;
        const uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        const uint32_t rot = (uint32_t)(state >> 59);
        const uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        for (int b = 0; b < 4; ++b) seed[4 * c + b] = (uint8_t)(x >> (8 * b));
    }
    pcg_from_seed(r, seed);
}

static double bits_to_double(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static uint64_t double_to_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }
// IntoFloat::into_float_with_exponent: 52 fraction bits, exponent e
static double with_exponent(uint64_t fraction, int e) { return bits_to_double(fraction | ((uint64_t)(1023 + e) << 52)); }

// ---- rand 0.8.5 Uniform<f64> (uniform.rs, UniformFloat)
typedef struct { double low, scale; } uniform_f64;

static void uniform_new(uniform_f64* u, double low, double high) {
    const double max_rand = with_exponent(UINT64_MAX >> 12, 0) - 1.0;
    double scale = high - low;
    while (scale * max_rand + low >= high) scale = bits_to_double(double_to_bits(scale) - 1);
    u->low = low;
    u->scale = scale;
}

static double uniform_sample(const uniform_f64* u, pcg64mcg* r) {
    const double value1_2 = with_exponent(pcg_next_u64(r) >> 12, 0);
    return (value1_2 - 1.0) * u->scale + u->low;
}

// ---- rand_distr 0.4.3 StandardNormal (normal.rs, utils.rs::ziggurat)
#define ZIG_R 3.654152885361008796
static double zig_x[257], zig_f[257];
static int zig_ready = 0;

static double literal18(double v) {  // the crate's table literals are the script's '%.18f' output
    char buf[64];
    snprintf(buf, sizeof buf, "%.18f", v);
    return strtod(buf, NULL);
}

static void zig_init(void) {
    if (zig_ready) return;
    const double r = 3.6541528853610088, v = 0.00492867323399;
    double x[257];
    x[0] = v / exp(-r * r / 2.0);
    x[1] = r;
    for (int i = 2; i < 256; ++i) {
        const double last = x[i - 1];
        x[i] = sqrt(-2.0 * log(v / last + exp(-last * last / 2.0)));
    }
    x[256] = 0.0;
    for (int i = 0; i <= 256; ++i) {
        zig_f[i] = literal18(exp(-x[i] * x[i] / 2.0));
        zig_x[i] = literal18(x[i]);
    }
    zig_ready = 1;
}

static double open01(pcg64mcg* r) { return with_exponent(pcg_next_u64(r) >> 12, 0) - (1.0 - 2.220446049250313e-16 / 2.0); }
static double standard_f64(pcg64mcg* r) { return (double)(pcg_next_u64(r) >> 11) * (1.0 / 9007199254740992.0); }

static double standard_normal(pcg64mcg* r) {
    for (;;) {
        const uint64_t bits = pcg_next_u64(r);
        const int i = (int)(bits & 0xff);
        const double u = with_exponent(bits >> 12, 1) - 3.0;
        const double x = u * zig_x[i];
        if (fabs(x) < zig_x[i + 1]) return x;
        if (i == 0) {  // the tail
            double xx = 1.0, yy = 0.0;
            while (-2.0 * yy < xx * xx) {
                const double x_ = open01(r), y_ = open01(r);
                xx = log(x_) / ZIG_R;
                yy = log(y_);
            }
            return u < 0.0 ? xx - ZIG_R : ZIG_R - xx;
        }
        if (zig_f[i + 1] + (zig_f[i] - zig_f[i + 1]) * standard_f64(r) < exp(-x * x / 2.0)) return x;
    }
}

// ---- exported
uint64_t refstream_pcg_new_next(uint64_t state_lo, uint64_t state_hi, uint64_t* out, int32_t count) {  // Mcg128Xsl64::new(state)
    pcg64mcg r;
    pcg_new(&r, (u128)state_lo | ((u128)state_hi << 64));
    for (int k = 0; k < count; ++k) out[k] = pcg_next_u64(&r);
    return count > 0 ? out[0] : 0;
}

void refstream_pcg_from_seed_next(const uint8_t* seed16, uint64_t* out, int32_t count) {
    pcg64mcg r;
    pcg_from_seed(&r, seed16);
    for (int k = 0; k < count; ++k) out[k] = pcg_next_u64(&r);
}

void refstream_pcg_seed_from_u64_next(uint64_t seed, uint64_t* out, int32_t count) {
    pcg64mcg r;
    pcg_seed_from_u64(&r, seed);
    for (int k = 0; k < count; ++k) out[k] = pcg_next_u64(&r);
}

void refstream_ziggurat_tables(double* x257, double* f257) {
    zig_init();
    memcpy(x257, zig_x, sizeof zig_x);
    memcpy(f257, zig_f, sizeof zig_f);
}

void refstream_normal(uint64_t seed, double mean, double std_dev, double* out, int64_t count) {
    zig_init();
    pcg64mcg r;
    pcg_seed_from_u64(&r, seed);
    for (int64_t k = 0; k < count; ++k) out[k] = mean + std_dev * standard_normal(&r);
}

void refstream_uniform(uint64_t seed, double low, double high, double* out, int64_t count) {
    pcg64mcg r;
    pcg_seed_from_u64(&r, seed);
    uniform_f64 u;
    uniform_new(&u, low, high);
    for (int64_t k = 0; k < count; ++k) out[k] = uniform_sample(&u, &r);
}

// The reference's generation loop (benchmark_target.rs:137-153) for n consecutive problems of `dof` DoFs, written SoA [dof][n]:
// nine arrays in the order current_position, current_velocity, current_acceleration, target_position, target_velocity,
// target_acceleration, max_velocity, max_acceleration, max_jerk. The three streams continue from problem to problem exactly as
// the reference's Randomizers do (a problem the reference skips after validate_input(false, false) has consumed its draws
// all the same; it is kept here and fails validation on every side alike).
void refstream_bench_inputs(int64_t n, int32_t dof, double** out9) {
    zig_init();
    pcg64mcg pos, dyn, lim;
    pcg_seed_from_u64(&pos, 42);
    pcg_seed_from_u64(&dyn, 43);
    pcg_seed_from_u64(&lim, 44);
    uniform_f64 u01, ulim;
    uniform_new(&u01, 0.0, 1.0);
    uniform_new(&ulim, 0.1, 12.0);
    const double p_keep[4] = {0.9, 0.8, 0.7, 0.6};
    for (int64_t i = 0; i < n; ++i) {
#define AT(a, d) out9[a][(int64_t)(d) * n + i]
        for (int d = 0; d < dof; ++d) AT(0, d) = 0.0 + 4.0 * standard_normal(&pos);
        for (int a = 0; a < 2; ++a)
            for (int d = 0; d < dof; ++d) AT(1 + a, d) = uniform_sample(&u01, &dyn) < p_keep[a] ? 0.0 + 0.8 * standard_normal(&dyn) : 0.0;
        for (int d = 0; d < dof; ++d) AT(3, d) = 0.0 + 4.0 * standard_normal(&pos);
        for (int a = 0; a < 2; ++a)
            for (int d = 0; d < dof; ++d) AT(4 + a, d) = uniform_sample(&u01, &dyn) < p_keep[2 + a] ? 0.0 + 0.8 * standard_normal(&dyn) : 0.0;
        for (int d = 0; d < dof; ++d) AT(6, d) = uniform_sample(&ulim, &lim) + fabs(AT(4, d));
        for (int d = 0; d < dof; ++d) AT(7, d) = uniform_sample(&ulim, &lim) + fabs(AT(5, d));
        for (int d = 0; d < dof; ++d) AT(8, d) = uniform_sample(&ulim, &lim);
#undef AT
    }
}
