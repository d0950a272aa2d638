// Host-side check of csrc/fftreg.cuh (compiled with g++ by tests/test_cpu_host.py): the register FFTs against a direct DFT,
// the pass-A / exchange / pass-B decomposition and the shuffle untangle emulated lane by lane against a direct real
// autocorrelation.  Prints "ok <max relative error>" or aborts.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../robust_speech_analysis_framework_b200/csrc/fftreg.cuh"

typedef std::complex<double> cd;
static const double PI = 3.14159265358979323846264338327950288;

template <int N, int SIGN>
static double check_fft() {
    double2 x[N];
    cd in[N];
    for (int i = 0; i < N; i++) { in[i] = cd(sin(1.3 * i + 0.2) + 0.1 * i, cos(0.7 * i * i)); x[i] = make_double2(in[i].real(), in[i].imag()); }
    fr_fft<N, SIGN>(x);
    int bits = 0; while ((1 << bits) < N) bits++;
    double err = 0, mag = 0;
    for (int k = 0; k < N; k++) {
        cd s = 0;
        for (int n = 0; n < N; n++) s += in[n] * std::polar(1.0, SIGN * 2 * PI * k * n / N);
        cd got(x[fr_brev(k, bits)].x, x[fr_brev(k, bits)].y);
        err = fmax(err, std::abs(got - s)); mag = fmax(mag, std::abs(s));
    }
    return err / mag;
}

// one frame, L lanes emulated in sequence; regs[j][slot]
template <int L, int SIGN>
static void transform(std::vector<std::vector<double2>>& regs) {
    const int M = 32 * L;
    std::vector<double2> xch(M);
    for (int j = 0; j < L; j++) {
        double2 a[32];
        for (int k = 0; k < 32; k++) a[k] = regs[j][k];
        fr_fft<32, SIGN>(a);
        for (int q = 0; q < 32; q++) {
            cd t = std::polar(1.0, SIGN * 2 * PI * (double)j * q / M);
            xch[fr_xch<L>(q, j)] = fr_mul(a[fr_brev(q, 5)], make_double2(t.real(), t.imag()));
        }
    }
    for (int j = 0; j < L; j++) {
        double2 b[32];
        if (L == 32) {
            for (int jj = 0; jj < 32; jj++) b[jj] = xch[fr_xch<L>(j, jj)];
            fr_fft<32, SIGN>(b);
        } else {
            for (int h = 0; h < 2; h++) {
                for (int jj = 0; jj < 16; jj++) b[16 * h + jj] = xch[fr_xch<L>(j + 16 * h, jj)];
                fr_fft<16, SIGN>(b + 16 * h);
            }
        }
        for (int s = 0; s < 32; s++) regs[j][s] = b[s];
    }
}

template <int L>
static double check_frame(int W) {
    const int M = 32 * L, N = 2 * M;
    std::vector<double> x(N, 0.0);
    for (int i = 0; i < W; i++) x[i] = (sin(0.05 * i) + 0.5 * sin(0.31 * i + 1) + 0.01 * ((i * 7919) % 13)) * (0.5 - 0.5 * cos(2 * PI * (i + 1) / (W + 1)));
    std::vector<std::vector<double2>> regs(L, std::vector<double2>(32));
    for (int j = 0; j < L; j++)
        for (int k = 0; k < 32; k++) { int n = j + L * k; regs[j][k] = make_double2(x[2 * n], x[2 * n + 1]); }
    transform<L, -1>(regs);
    // untangle, lock-step over r like the shuffle code
    for (int r = 0; r < 16; r++) {
        std::vector<double2> send(L), yk(L), ymk(L);
        for (int j = 0; j < L; j++) send[j] = regs[j][fr_slot<L>(31 - r)];
        for (int j = 0; j < L; j++) {
            int pl = (L - j) % L;
            double2 theirs = send[pl], mine = regs[j][fr_slot<L>(r)];
            if (j == 0 && r >= 1) theirs = regs[j][fr_slot<L>(32 - r)];
            cd w = std::polar(1.0, -2 * PI * j / N) * cd(fr_cos64(r), -fr_sin64(r));
            double pk, pmk;
            fr_pair(mine, theirs, make_double2(w.real(), w.imag()), &yk[j], &ymk[j], &pk, &pmk);
            if (j == 0 && r == 0) {
                double p0 = (mine.x + mine.y) * (mine.x + mine.y), pM = (mine.x - mine.y) * (mine.x - mine.y);
                yk[j] = make_double2(p0 + pM, p0 - pM);
            }
        }
        for (int j = 0; j < L; j++) {
            int pl = (L - j) % L;
            regs[j][fr_slot<L>(r)] = yk[j];
            if (j != 0) regs[j][fr_slot<L>(31 - r)] = ymk[pl];
            else if (r >= 1) regs[j][fr_slot<L>(32 - r)] = ymk[j];
        }
    }
    {
        double2 z = regs[0][fr_slot<L>(16)], yk, ymk; double pk, pmk;
        cd w = std::polar(1.0, -2 * PI * (M / 2) / N);
        fr_pair(z, z, make_double2(w.real(), w.imag()), &yk, &ymk, &pk, &pmk);
        regs[0][fr_slot<L>(16)] = yk;
    }
    // logical order for the inverse
    std::vector<std::vector<double2>> in2(L, std::vector<double2>(32));
    for (int j = 0; j < L; j++) for (int r = 0; r < 32; r++) in2[j][r] = regs[j][fr_slot<L>(r)];
    transform<L, +1>(in2);
    std::vector<double> ac(N);
    for (int j = 0; j < L; j++) for (int r = 0; r < 32; r++) { int n = j + L * r; ac[2 * n] = in2[j][fr_slot<L>(r)].x; ac[2 * n + 1] = in2[j][fr_slot<L>(r)].y; }
    double err = 0;
    std::vector<double> ref(N / 2);
    for (int k = 0; k < N / 2; k++) { double s = 0; for (int i = 0; i + k < W; i++) s += x[i] * x[i + k]; ref[k] = s; }
    for (int k = 0; k <= W / 2; k++) err = fmax(err, fabs(ac[k] / ac[0] - ref[k] / ref[0]));
    return err;
}

int main() {
    double e = 0;
    e = fmax(e, check_fft<8, -1>()); e = fmax(e, check_fft<16, -1>()); e = fmax(e, check_fft<32, -1>());
    e = fmax(e, check_fft<16, 1>()); e = fmax(e, check_fft<32, 1>());
    e = fmax(e, check_frame<16>(478)); e = fmax(e, check_frame<16>(638));
    e = fmax(e, check_frame<32>(958)); e = fmax(e, check_frame<32>(798));
    if (!(e < 1e-13)) { printf("FAIL %g\n", e); return 1; }
    printf("ok %.3g\n", e);
    return 0;
}
