/*
 * binary_io.c -- "-b write" / "-b read": save and reload the generated problem.
 *
 * The reference dumps the SimulationData struct verbatim (stale pointers included) followed
 * by the six arrays, with no header and unchecked freads (cuda/io.cu:443-495); a file from
 * one port cannot be read by another because the struct differs.  This format keeps the
 * same file name and array order but starts with a validated header.
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

SimulationData binary_read(Inputs in)
{
    printf("Reading all data structures from binary file %s...\n", XS_FILE_NAME);
    FILE *fp = fopen(XS_FILE_NAME, "rb");
    if (!fp) die("cannot open for reading");
    xs_file_header h;
    if (fread(&h, sizeof h, 1, fp) != 1) die("truncated header");
    if (h.magic != XS_FILE_MAGIC || h.format_version != 1) die("not an xsbench_b200 data file");
    if (h.grid_type != in.grid_type || h.n_isotopes != in.n_isotopes ||
        h.n_gridpoints != in.n_gridpoints ||
        (in.grid_type == XS_HASH && h.hash_bins != in.hash_bins))
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
