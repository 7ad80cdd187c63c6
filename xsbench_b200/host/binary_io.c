/*
 * binary_io.c -- "-b write" / "-b read": save and reload the generated problem.
 *
 * The reference dumps the SimulationData struct verbatim (stale pointers included) followed
 * by the six arrays, with no header and unchecked freads (cuda/io.cu:443-495); a file from
 * one port cannot be read by another because the struct differs.  This format keeps the
 * same file name and array order but starts with a validated header.
 *
 * binary_read also accepts the reference's own files: when the magic is absent the file is taken
 * as a raw struct dump -- the cuda/ port's 128-byte SimulationData or the openmp-threading port's
 * 112-byte one (the two share their first 80 bytes: six stale pointers, then the lengths) --
 * followed by the six arrays in the same order; which of the two it is follows from the file size,
 * and every length is checked against the problem the command line asks for.
 */
#include "xs_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define XS_FILE_NAME  "XS_data.dat"
#define XS_FILE_MAGIC 0x3030324253585358ULL   /* "XSXSB200" little-endian-ish tag */

typedef struct {
    uint64_t magic;
    int32_t  format_version;
    int32_t  grid_type;
    int64_t  n_isotopes, n_gridpoints;
    int32_t  hash_bins, max_num_nucs;
    int64_t  len_num_nucs, len_concs, len_mats, len_nuclide_grid, len_index_grid, len_ueg;
} xs_file_header;

static void die(const char *what)
{
    fprintf(stderr, "binary file %s: %s\n", XS_FILE_NAME, what);
    exit(1);
}

void binary_write(Inputs in, SimulationData SD)
{
    printf("Writing all data structures to binary file %s...\n", XS_FILE_NAME);
    FILE *fp = fopen(XS_FILE_NAME, "wb");
    if (!fp) die("cannot open for writing");
    xs_file_header h = { XS_FILE_MAGIC, 1, in.grid_type, in.n_isotopes, in.n_gridpoints,
                         in.hash_bins, SD.max_num_nucs, SD.length_num_nucs, SD.length_concs,
                         SD.length_mats, SD.length_nuclide_grid, SD.length_index_grid,
                         SD.length_unionized_energy_array };
    int ok = fwrite(&h, sizeof h, 1, fp) == 1;
    ok &= fwrite(SD.num_nucs, sizeof(int), (size_t)h.len_num_nucs, fp) == (size_t)h.len_num_nucs;
    ok &= fwrite(SD.concs, sizeof(double), (size_t)h.len_concs, fp) == (size_t)h.len_concs;
    ok &= fwrite(SD.mats, sizeof(int), (size_t)h.len_mats, fp) == (size_t)h.len_mats;
    ok &= fwrite(SD.nuclide_grid, sizeof(NuclideGridPoint), (size_t)h.len_nuclide_grid, fp)
              == (size_t)h.len_nuclide_grid;
    ok &= fwrite(SD.index_grid, sizeof(int), (size_t)h.len_index_grid, fp) == (size_t)h.len_index_grid;
    ok &= fwrite(SD.unionized_energy_array, sizeof(double), (size_t)h.len_ueg, fp) == (size_t)h.len_ueg;
    if (fclose(fp) != 0 || !ok) die("short write");
}

static void *read_array(FILE *fp, size_t elem, int64_t count)
{
    if (count < 0) die("negative length");
    void *p = malloc(elem * (size_t)(count ? count : 1));
    if (!p) die("out of memory");
    if (fread(p, elem, (size_t)count, fp) != (size_t)count) die("truncated");
    return p;
}

/* The prefix the reference's cuda/ and openmp-threading/ SimulationData share
 * (cuda/XSbench_header.cuh:65-85, openmp-threading/XSbench_header.h:77-102). */
typedef struct {
    uint64_t stale_pointers[6];
    int32_t  length_num_nucs, length_concs, length_mats, length_unionized_energy_array;
    int64_t  length_index_grid;
    int32_t  length_nuclide_grid, max_num_nucs;
} xs_legacy_prefix;

static long file_size(FILE *fp)
{
    long here = ftell(fp), size = -1;
    if (fseek(fp, 0, SEEK_END) == 0) size = ftell(fp);
    fseek(fp, here, SEEK_SET);
    return size;
}

SimulationData binary_read(Inputs in)
{
    printf("Reading all data structures from binary file %s...\n", XS_FILE_NAME);
    FILE *fp = fopen(XS_FILE_NAME, "rb");
    if (!fp) die("cannot open for reading");
    xs_file_header h;
    if (fread(&h, sizeof h, 1, fp) != 1) die("truncated header");
    if (h.magic != XS_FILE_MAGIC) {
        /* a file written by the reference itself (cuda/io.cu:443-462): raw struct, then the arrays */
        xs_legacy_prefix lp;
        memcpy(&lp, &h, sizeof lp);
        memset(&h, 0, sizeof h);
        h.format_version = 1;
        h.grid_type = in.grid_type; h.n_isotopes = in.n_isotopes; h.n_gridpoints = in.n_gridpoints;
        h.hash_bins = in.hash_bins; h.max_num_nucs = lp.max_num_nucs;
        h.len_num_nucs = lp.length_num_nucs; h.len_concs = lp.length_concs; h.len_mats = lp.length_mats;
        h.len_nuclide_grid = lp.length_nuclide_grid; h.len_index_grid = lp.length_index_grid;
        h.len_ueg = lp.length_unionized_energy_array;
        if (h.len_num_nucs != XS_NUM_MATERIALS || h.max_num_nucs < 1 ||
            h.len_concs != (int64_t)XS_NUM_MATERIALS * h.max_num_nucs || h.len_mats != h.len_concs ||
            h.len_index_grid < 0 || h.len_ueg < 0)
            die("neither an xsbench_b200 data file nor a reference struct dump");
        const long payload = (long)(h.len_num_nucs * 4 + h.len_concs * 8 + h.len_mats * 4 + h.len_nuclide_grid * 48
                                    + h.len_index_grid * 4 + h.len_ueg * 8);
        const long struct_bytes = file_size(fp) - payload;
        if (struct_bytes != 128 && struct_bytes != 112)
            die("neither an xsbench_b200 data file nor a reference struct dump");
        printf("(reference-format file: %ld-byte SimulationData dump of the %s port)\n", struct_bytes,
               struct_bytes == 128 ? "cuda" : "openmp-threading");
        if (fseek(fp, struct_bytes, SEEK_SET) != 0) die("truncated");
    } else {
        if (h.format_version != 1) die("unknown format version");
        if (h.grid_type != in.grid_type || (in.grid_type == XS_HASH && h.hash_bins != in.hash_bins))
            die("file was written for a different problem (-s/-g/-G/-h)");
    }
    /* either format: the lengths must be the ones this command line implies */
    const int64_t n_points = (int64_t)in.n_isotopes * in.n_gridpoints;
    const int64_t want_ueg = in.grid_type == XS_UNIONIZED ? n_points : 0;
    const int64_t want_index = in.grid_type == XS_UNIONIZED ? n_points * in.n_isotopes
                             : in.grid_type == XS_HASH ? (int64_t)in.hash_bins * in.n_isotopes : 0;
    if (h.n_isotopes != in.n_isotopes || h.n_gridpoints != in.n_gridpoints || h.len_nuclide_grid != n_points ||
        h.len_ueg != want_ueg || h.len_index_grid != want_index || h.len_num_nucs != XS_NUM_MATERIALS ||
        h.max_num_nucs < 1 || h.len_concs != (int64_t)XS_NUM_MATERIALS * h.max_num_nucs || h.len_mats != h.len_concs)
        die("file was written for a different problem (-s/-g/-G/-h)");

    SimulationData SD;
    memset(&SD, 0, sizeof SD);
    SD.length_num_nucs = (int)h.len_num_nucs;
    SD.length_concs = (int)h.len_concs;
    SD.length_mats = (int)h.len_mats;
    SD.length_nuclide_grid = (int)h.len_nuclide_grid;
    SD.length_index_grid = (long)h.len_index_grid;
    SD.length_unionized_energy_array = (int)h.len_ueg;
    SD.max_num_nucs = h.max_num_nucs;
    SD.num_nucs = (int *)read_array(fp, sizeof(int), h.len_num_nucs);
    SD.concs = (double *)read_array(fp, sizeof(double), h.len_concs);
    SD.mats = (int *)read_array(fp, sizeof(int), h.len_mats);
    SD.nuclide_grid = (NuclideGridPoint *)read_array(fp, sizeof(NuclideGridPoint), h.len_nuclide_grid);
    SD.index_grid = (int *)read_array(fp, sizeof(int), h.len_index_grid);
    SD.unionized_energy_array = (double *)read_array(fp, sizeof(double), h.len_ueg);
    fclose(fp);
    return SD;
}
