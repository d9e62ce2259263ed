/*
 * bs_io.h -- C ABI of the blackscholes loader and writer (part of libbs_gpu.so).
 *
 * Replaces, with identical file grammar and identical resulting values:
 *   loader  /root/reference/parsec-ff/pkgs/apps/blackscholes/src/blackscholes.c:696-739
 *           (header fscanf("%i"), then N x fscanf("%f %f %f %f %f %f %c %f %f"), rv != 9 -> error)
 *           fused with the AoS->SoA staging of :760-767 (otype = (OptionType == 'P') ? 1 : 0):
 *           the rows are parsed by all host cores straight into the caller's SoA arrays -- normally the
 *           pinned buffers of bs_gpu_host_buffer() -- so no AoS copy ever exists.
 *   writer  blackscholes.c:923-947  ("%i\n" then N x "%.18f\n"), formatted by all host cores.
 * Values are bit-identical to what fscanf/fprintf produce (both sides round correctly); input that is
 * not plain whitespace-separated tokens is re-read with the C library's own fscanf so that even odd
 * files behave exactly as under the reference.
 */
#ifndef BS_IO_H
#define BS_IO_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bs_io_status {
    BS_IO_OK = 0,
    BS_IO_ERR_OPEN = -1,   /* reference: "ERROR: Unable to open file `%s'."       (:697-700, :924-928) */
    BS_IO_ERR_READ = -2,   /* reference: "ERROR: Unable to read from file `%s'."  (:702-706, :729-733) */
    BS_IO_ERR_WRITE = -3,  /* reference: "ERROR: Unable to write to file `%s'."   (:930-934, :937-941) */
    BS_IO_ERR_CLOSE = -4,  /* reference: "ERROR: Unable to close file `%s'."      (:735-739, :943-947) */
    BS_IO_ERR_INVALID = -5,
    BS_IO_ERR_NOMEM = -6
} bs_io_status;

typedef struct bs_io_file bs_io_file;

/* Open an input file and read its header.  *num_options receives the "%i" value (which, like the
 * reference, may be written in decimal, 0x-hex or 0-octal). */
int bs_io_open(const char *path, bs_io_file **file, long long *num_options);

/* Parse the first `count` rows into SoA arrays of `fp_bytes`-wide fptype (4 or 8; 8 parses with %lf
 * semantics).  dgrefval/divq/divs may be NULL.  nthreads <= 0 means all host cores.
 * Fails with BS_IO_ERR_READ if fewer than `count` complete rows are present. */
int bs_io_load(bs_io_file *file, int fp_bytes, size_t count, void *sptprice, void *strike, void *rate,
               void *volatility, void *otime, int *otype, void *dgrefval, void *divq, void *divs, int nthreads);

/* ---- binary SoA side-car -------------------------------------------------------
 * A ".bssoa" file holds the already parsed streams (4 KiB header, then sptprice, strike, rate, volatility, otime,
 * otype, DGrefval, each padded to 4 KiB) so that repeat runs skip the text parse and a 1B-option set fits on disk
 * (28 GB instead of ~62 GB of text).  bs_io_open() recognises the format by its magic: bs_io_load() then copies the
 * streams instead of parsing (the file's fptype must equal fp_bytes, else BS_IO_ERR_READ).  A side-car written with
 * `source_path` records that file's size and mtime; bs_io_soa_matches() tells whether it is still current. */
int bs_io_soa_write(const char *path, int fp_bytes, size_t count, const void *sptprice, const void *strike,
                    const void *rate, const void *volatility, const void *otime, const int *otype,
                    const void *dgrefval, const char *source_path);
/* 1 if `soa_path` is a side-car of `source_path` (same size and mtime) holding fp_bytes-wide streams, else 0. */
int bs_io_soa_matches(const char *soa_path, const char *source_path, int fp_bytes);
/* 1 if the opened file is a binary side-car, 0 if it is text. */
int bs_io_is_soa(const bs_io_file *file);

/* Close the input file (the reference's fclose at :735). */
int bs_io_close(bs_io_file *file);

/* Write the prices file. */
int bs_io_write_prices(const char *path, int fp_bytes, size_t count, const void *prices, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* BS_IO_H */
