// Prints detail::mul44 / detail::inv_affine of include/cvo.hpp on 4x4 float matrices read from stdin
// (32 floats per case: a then b) as raw bit patterns, for the CPU test that pins the Python mirror
// (cvo_slam_b200/cvo.py: _mul44_f32, _inv_affine_f32) to the drop-in header bit for bit.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "cvo.hpp"

int main() {
    float a[16], b[16], c[16], inv[16];
    for (;;) {
        for (int i = 0; i < 16; i++) if (scanf("%f", &a[i]) != 1) return 0;
        for (int i = 0; i < 16; i++) if (scanf("%f", &b[i]) != 1) return 0;
        cvo::detail::mul44(a, b, c);
        cvo::detail::inv_affine(c, inv);
        for (int i = 0; i < 16; i++) { uint32_t u; memcpy(&u, &c[i], 4); printf("%08x ", u); }
        for (int i = 0; i < 16; i++) { uint32_t u; memcpy(&u, &inv[i], 4); printf("%08x ", u); }
        printf("\n");
    }
}
