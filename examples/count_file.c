/* Minimal C client of the drop-in boundary: count the k-mers of one FASTA file and write the per-sample TSV, the way
 * run_mercat2 does for one sample (reference bin/mercat2.py:115-137).
 *
 *   gcc -std=c99 -Iinclude examples/count_file.c -Lmercat2_b200 -lmercat2_b200 -Wl,-rpath,$PWD/mercat2_b200 -o count_file
 *   ./count_file sample.fna 31 10 100 sample_counts.tsv
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "mercat2_b200.h"

static int fail(const char* what, int rc) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, mc2_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s <fasta[.gz]> <k> <min_count> <chunk MB, 0 = off> <out.tsv>\n", argv[0]);
        return 2;
    }
    const char* path = argv[1];
    const int k = atoi(argv[2]);
    const long long min_count = atoll(argv[3]);
    const unsigned long long chunk_mb = strtoull(argv[4], NULL, 10);
    struct stat st;
    if (stat(path, &st) != 0) { perror(path); return 1; }
    /* the reference chunks only when the ON-DISK size reaches -s MiB (bin/mercat2.py:101) */
    const uint64_t chunk_bytes = (chunk_mb && (uint64_t)st.st_size >= chunk_mb * 1024 * 1024) ? chunk_mb * 1024 * 1024 : 0;

    mc2_engine* engine = NULL;
    mc2_sample* sample = NULL;
    mc2_table* table = NULL;
    int rc = mc2_engine_create(0, &engine);
    if (rc) return fail("mc2_engine_create", rc);
    rc = mc2_sample_begin(engine, k, min_count, &sample);
    if (rc) return fail("mc2_sample_begin", rc);
    uint64_t pieces = 0, text_bytes = 0;
    rc = mc2_sample_add_file(sample, path, -1, chunk_bytes, &pieces, &text_bytes);
    if (rc) { mc2_sample_abort(sample); return fail("mc2_sample_add_file", rc); }
    rc = mc2_sample_finish(sample, &table);                 /* consumes the sample */
    if (rc) return fail("mc2_sample_finish", rc);
    const uint64_t rows = mc2_table_rows(table);
    if (rows) {
        printf("Significant k-mers: %llu\n", (unsigned long long)rows);
        rc = mc2_table_write_tsv(table, argv[5], "sample");
        if (rc < 0) return fail("mc2_table_write_tsv", rc);
    } else {
        printf("No significant k-mers found\n");
    }
    fprintf(stderr, "%llu bytes of text in %llu piece(s)\n", (unsigned long long)text_bytes, (unsigned long long)pieces);
    mc2_table_free(table);
    mc2_engine_destroy(engine);
    return 0;
}
