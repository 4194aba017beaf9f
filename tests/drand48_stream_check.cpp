// Drand48Stream (host/include/gmapping/utils/stat.h) against libc: same Gaussian samples bit for bit, and drand48() continues where a
// stream stopped.  Built and run by tests/test_host_layer.py::test_drand48_stream_matches_libc.
#include <gmapping/utils/stat.h>
#include <cstdio>
#include <cstring>
using namespace GMapping;
int main() {
    srand48(12345);
    double ref[2000]; for (int i = 0; i < 2000; i++) ref[i] = sampleGaussian(0.01 + i * 1e-4);
    double after_ref = drand48();
    srand48(12345);
    double got[2000];
    { Drand48Stream s; for (int i = 0; i < 1000; i++) got[i] = s.gaussian(0.01 + i * 1e-4); }
    for (int i = 1000; i < 1500; i++) got[i] = sampleGaussian(0.01 + i * 1e-4);   // libc continues where the stream stopped
    { Drand48Stream s; for (int i = 1500; i < 2000; i++) got[i] = s.gaussian(0.01 + i * 1e-4); }
    double after_got = drand48();
    int bad = 0; for (int i = 0; i < 2000; i++) bad += memcmp(&ref[i], &got[i], 8) != 0;
    printf("mismatches %d, next draw equal %d\n", bad, after_ref == after_got);
    // plain uniform draws
    srand48(7); double u1[100]; for (int i = 0; i < 100; i++) u1[i] = drand48();
    srand48(7); int b2 = 0; { Drand48Stream s; for (int i = 0; i < 100; i++) b2 += s.next() != u1[i]; }
    printf("uniform mismatches %d\n", b2);
    return bad || b2 || after_ref != after_got;
}
