// Reads lines "op a0 a1 a2 a3 b0 b1 b2 b3" (hex limbs, Montgomery form irrelevant: the routines are plain modular arithmetic on
// limbs with a Montgomery product) and prints the result limbs of halo2::fr::{mul, add, neg(a), from_u64(a0)}.  CPU only.
#include <cinttypes>
#include <cstdio>

#include "zkb200_halo2.hpp"

int main() {
    char op[8];
    unsigned long long v[8];
    while (std::scanf("%7s %llx %llx %llx %llx %llx %llx %llx %llx", op, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]) == 9) {
        const halo2::Fr a = {v[0], v[1], v[2], v[3]}, b = {v[4], v[5], v[6], v[7]};
        halo2::Fr r{};
        if (op[0] == 'm') r = halo2::fr::mul(a, b);
        else if (op[0] == 'a') r = halo2::fr::add(a, b);
        else if (op[0] == 'n') r = halo2::fr::neg(a);
        else r = halo2::fr::from_u64(a[0]);
        std::printf("%016" PRIx64 " %016" PRIx64 " %016" PRIx64 " %016" PRIx64 "\n", r[0], r[1], r[2], r[3]);
    }
    return 0;
}
