/*
 * bs_inputgen.c -- synthetic blackscholes input files in PARSEC inputgen format.
 *
 *   bs_inputgen <numOptions> <fileName>
 *
 * Writes the header line "<numOptions>\n" followed by numOptions rows
 *   "%.2f %.2f %.4f %.2f %.2f %.2f %c %.2f %.18f\n"   (s strike r divq v t type divs DGrefval)
 * where row i is bs_option_table[i % 1000] -- the grammar the reference loader consumes with
 * fscanf("%i") + fscanf("%f %f %f %f %f %f %c %f %f")
 * (parsec-ff/pkgs/apps/blackscholes/src/blackscholes.c:701,728).  The 1000 rows are formatted once
 * and then replayed, so the 10M-row native file (about 620 MB) takes a second or two to write.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bs_option_table.h"

#define ROW_MAX 160

int main(int argc, char **argv)
{
    static char rows[BS_TABLE_ROWS][ROW_MAX];
    static int row_len[BS_TABLE_ROWS];
    long long n, i;
    FILE *f;
    char *chunk;
    size_t chunk_cap = 8u << 20, fill = 0;

    if (argc != 3) {
        printf("Usage:\n\t%s <numOptions> <fileName>\n", argv[0]);
        return 1;
    }
    n = atoll(argv[1]);
    if (n < 0 || n > 2147483647LL) {
        printf("ERROR: numOptions must fit the reference's `int numOptions`.\n");
        return 1;
    }
    f = fopen(argv[2], "w");
    if (!f) {
        printf("ERROR: Unable to open file `%s'.\n", argv[2]);
        return 1;
    }
    for (i = 0; i < BS_TABLE_ROWS; i++) {
        const bs_table_row *r = &bs_option_table[i];
        row_len[i] = snprintf(rows[i], ROW_MAX, "%.2f %.2f %.4f %.2f %.2f %.2f %c %.2f %.18f\n",
                              r->s, r->strike, r->r, r->divq, r->v, r->t, r->option_type, r->divs,
                              r->dgrefval);
    }
    chunk = (char *)malloc(chunk_cap + ROW_MAX);
    if (!chunk) { fclose(f); return 1; }
    fprintf(f, "%lld\n", n);
    for (i = 0; i < n; i++) {
        int k = (int)(i % BS_TABLE_ROWS);
        memcpy(chunk + fill, rows[k], (size_t)row_len[k]);
        fill += (size_t)row_len[k];
        if (fill >= chunk_cap) {
            if (fwrite(chunk, 1, fill, f) != fill) { printf("ERROR: short write.\n"); return 1; }
            fill = 0;
        }
    }
    if (fill && fwrite(chunk, 1, fill, f) != fill) { printf("ERROR: short write.\n"); return 1; }
    free(chunk);
    if (fclose(f) != 0) {
        printf("ERROR: Unable to close file `%s'.\n", argv[2]);
        return 1;
    }
    return 0;
}
