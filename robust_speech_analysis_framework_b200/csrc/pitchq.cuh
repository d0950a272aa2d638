// pitchq.cuh -- Pitch queries on the selected path (fon/Sampled.cpp Sampled_getValueAtX with Pitch::v_getValueAtSample,
// Hertz, linear): pitch.get_value_at_time at mshds_extractor.py:109,364 and Pitch_getValueAtTime inside
// Sound_Pitch_to_PointProcess_cc.
#pragma once
#include "common.cuh"

struct PitchView {
    const double* f;     // selected frequency per frame (0-based storage of frames 1..nx)
    int nx;
    double x1, dx, ceiling, xmin, xmax;
};

__device__ __forceinline__ bool pv_voiced(const PitchView& p, long long i /*1-based*/) {
    if (i < 1 || i > p.nx) return false;
    double f = p.f[i - 1];
    return f > 0.0 && f < p.ceiling;
}

__device__ __forceinline__ double pitch_value_at(const PitchView& p, double x) {
    if (x < p.xmin || x > p.xmax) return DEVNAN;
    double ireal = (x - p.x1) / p.dx + 1.0;
    long long ileft = (long long)floor(ireal), inear, ifar;
    double phase = ireal - (double)ileft;
    if (phase < 0.5) { inear = ileft; ifar = ileft + 1; }
    else { ifar = ileft; inear = ileft + 1; phase = 1.0 - phase; }
    if (inear < 1 || inear > p.nx) return DEVNAN;
    if (!pv_voiced(p, inear)) return DEVNAN;
    double fnear = p.f[inear - 1];
    if (ifar < 1 || ifar > p.nx) return fnear;
    if (!pv_voiced(p, ifar)) return fnear;
    double ffar = p.f[ifar - 1];
    return fnear + phase * (ffar - fnear);
}
