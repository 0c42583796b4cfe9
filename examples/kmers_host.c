/* Plain-C caller of libkameris_b200.so: what a rebuilt generation_cgr (kameris-backend) would do in place
 * of its per-file lambda + ParallelExecutor -- read record[0] of every file named on the command line and
 * print the k-mer count vectors' checksums.
 *
 *   gcc -std=c99 -Iinclude examples/kmers_host.c -Lkameris_b200 -lkameris_b200 -Wl,-rpath,$PWD/kameris_b200 -o kmers_host
 *   ./kmers_host 6 a.fasta b.fasta
 */
#include <stdio.h>
#include <stdlib.h>

#include "kameris_b200.h"

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <k> <fasta files...>\n", argv[0]);
        return 1;
    }
    const int k = atoi(argv[1]);
    const uint32_t n = (uint32_t)(argc - 2);
    const char *const *paths = (const char *const *)(argv + 2);
    uint64_t *sizes = (uint64_t *)calloc(n, sizeof *sizes);
    uint64_t *offs = (uint64_t *)calloc(n + 1, sizeof *offs);
    uint64_t *start = (uint64_t *)calloc(n + 1, sizeof *start);
    uint32_t bad = 0, i;
    if (kam_fasta_file_sizes(paths, n, sizes, 0)) goto fail;
    for (i = 0; i < n; ++i) offs[i + 1] = offs[i] + sizes[i];
    {
        uint8_t *chars = (uint8_t *)malloc(offs[n] + 16);
        const size_t bins = (size_t)1 << (2 * k);
        uint16_t *counts = (uint16_t *)malloc((size_t)n * bins * sizeof *counts);
        if (kam_fasta_read_files(paths, n, offs, chars, start, 0, &bad)) goto fail;
        if (kam_kmers_host(chars, start, n, k, 16, KAM_OUT_COUNTS, counts)) goto fail;
        for (i = 0; i < n; ++i) {
            unsigned long long sum = 0;
            size_t b;
            for (b = 0; b < bins; ++b) sum += counts[(size_t)i * bins + b];
            printf("%s\t%llu bases\t%llu k-mers\n", argv[2 + i], (unsigned long long)(start[i + 1] - start[i]), sum);
        }
        free(counts);
        free(chars);
    }
    free(sizes);
    free(offs);
    free(start);
    return 0;
fail:
    fprintf(stderr, "error: %s\n", kam_last_error());
    return 1;
}
