// CPU run of the Gadget-scheme initial-condition generator of fastpm_b200/csrc/ic_gadget.h (the function the CUDA kernel
// calls per (kx, ky) column): writes delta_k for an n^3 mesh in the reference's untransposed layout [kx][ky][kz] complex64 to
// the file given on the command line.  tests/test_cpu_oracle_and_host.py compares it bit for bit with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fastpm_b200/csrc/ic_gadget.h"

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: ic_emul n seed out.bin\n"); return 2; }
    const int n = atoi(argv[1]), seed = atoi(argv[2]);
    std::vector<unsigned int> self((size_t) n * n), conj((size_t) n * n);
    fpm_gadget_seed_table(n, seed, self.data(), conj.data());
    const int hc = n / 2 + 1;
    std::vector<FpmFloat2> out((size_t) n * n * hc);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            fpm_gadget_fill_column(n, i, j, self[(size_t) i * n + j], conj[(size_t) i * n + j], out.data() + ((size_t) i * n + j) * hc);
    FILE *f = fopen(argv[3], "wb");
    if (!f) return 3;
    fwrite(out.data(), sizeof(FpmFloat2), out.size(), f);
    fclose(f);
    return 0;
}
