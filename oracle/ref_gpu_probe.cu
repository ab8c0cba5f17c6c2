/* oracle/ref_gpu_probe.cu -- TEST / BENCH INFRASTRUCTURE ONLY.
 *
 * A ~60-line driver (ours) around the UNMODIFIED reference: oracle/build_ref.sh compiles it with
 * -I/root/reference so that the `#include "core.cu"` below pulls the reference's whole translation
 * unit from where it lies (exactly as its own main.cu:4 does); nothing of the reference is stored in
 * this repository.  It runs ONE variant on ONE shape, on the data the reference's own driver would
 * generate (main.cu:10-13, 24-35, 54, 64: srand(1000); rand()/double(RAND_MAX); queries first), timed
 * by wall clock around the whole callback like main.cu:73-76 (malloc + H2D + kernel + D2H inside),
 * and writes the returned indices to a file so that bench.py can compare them with V0.
 *
 *     ref_gpu_probe <variant 0..9> <k> <m> <n> <reps> <indices.bin>
 *
 * Output: one line  "ref_gpu variant=V k=K m=M n=N reps=R best_ms=... median_ms=..."
 * Used by bench.py's `ref_gpu` side record (the reference's best valid GPU kernel, V7
 * core.cu:589-633, beside this engine's number on the same box). */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>
#include <vector>
#include "core.cu"

typedef void (*call_fn)(int, int, int, float *, float *, int **);

int main(int argc, char **argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: %s variant k m n reps out.bin\n", argv[0]);
        return 2;
    }
    static const call_fn table[10] = {v0::cudaCall, v1::cudaCall, v2::cudaCall, v3::cudaCall, v4::cudaCall,
                                      v5::cudaCall, v6::cudaCall, v7::cudaCall, v8::cudaCall, v9::cudaCall};
    const int v = atoi(argv[1]), k = atoi(argv[2]), m = atoi(argv[3]), n = atoi(argv[4]), reps = atoi(argv[5]);
    if (v < 0 || v > 9 || k <= 0 || m <= 0 || n <= 0 || reps <= 0) return 2;
    srand(1000);
    float *s = (float *)malloc(sizeof(float) * (size_t)k * m), *r = (float *)malloc(sizeof(float) * (size_t)k * n);
    for (size_t i = 0; i < (size_t)k * m; i++) s[i] = rand() / double(RAND_MAX);
    for (size_t i = 0; i < (size_t)k * n; i++) r[i] = rand() / double(RAND_MAX);
    std::vector<double> ms;
    int *results = NULL;
    for (int it = 0; it < reps + 1; ++it) {  // first call = warm-up (not reported)
        if (results) free(results);
        results = NULL;
        const long t0 = getTime();
        table[v](k, m, n, s, r, &results);
        const long t1 = getTime();
        if (it > 0) ms.push_back((t1 - t0) / 1e6);
    }
    std::sort(ms.begin(), ms.end());
    FILE *f = fopen(argv[6], "wb");
    if (f) {
        fwrite(results, sizeof(int), (size_t)m, f);
        fclose(f);
    }
    printf("ref_gpu variant=%d k=%d m=%d n=%d reps=%d best_ms=%.4f median_ms=%.4f\n", v, k, m, n, reps, ms.front(),
           ms[ms.size() / 2]);
    return 0;
}
