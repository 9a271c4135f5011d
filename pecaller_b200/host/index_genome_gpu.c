/*
 * index_genome_gpu.c - writes the four index files of index_genome_whole (wingolab-org/pecaller
 * src/index_genome_whole.c main(), 93-354) with the 16-mer index built on a B200 through the C-ABI
 * (pemap_init_from_genome): G.sdx (text), G.seq (gz of the upper-cased letters), G.idx (gz of the 2^32+1 prefix
 * table), G.mdx (raw positions).  The inflated contents are byte-identical to the reference's; the reference needs
 * ~3.3 minutes and 64 GiB of virtual memory for any genome, the device build about a second plus the file writes.
 *
 *   index_genome_gpu genome.fa basename [y|n]          (bisulfite index: y)
 * or, without arguments, the reference's interactive prompts on stdin
 * ([S,D] [log file] max_contigs fasta basename bisulfite; index_genome_whole.c:117-166).
 * Written from scratch; what must be byte-compatible cites the reference line.
 */
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "pemap.h"

static void die(const char *msg) {
  printf("\n%s\n", msg);
  exit(1);
}

static void read_var(const char *prompt, char *dst, size_t cap) { /* read_var, 860-879 */
  printf("%s", prompt);
  if (!fgets(dst, (int)cap, stdin)) die(" Unexpected end of input ");
  dst[strcspn(dst, "\r\n")] = '\0';
}

int main(int argc, char **argv) {
  char fasta[1024], base[1024], ans[1024];
  int bisulfite = 0;
  if (argc >= 3) {
    strcpy(fasta, argv[1]);
    strcpy(base, argv[2]);
    bisulfite = argc > 3 && (strchr(argv[3], 'Y') || strchr(argv[3], 'y'));
  } else {
    read_var("\nSend Output to Screen or Disk? [S,D]\n", ans, sizeof ans);
    if (strchr(ans, 'D') || strchr(ans, 'd')) read_var("Please Enter File Name for Output\n", ans, sizeof ans);
    read_var("Maximum Number of Contig Fasta Files to Process\n", ans, sizeof ans);
    read_var("Please Enter Name For Fastaq File\n", fasta, sizeof fasta);
    read_var("Basename to save compressed Genome and Indexes\n", base, sizeof base);
    read_var("Will the target DNA be bisulfite converted?\n", ans, sizeof ans);
    bisulfite = strchr(ans, 'Y') || strchr(ans, 'y');
  }
  FILE *in = fopen(fasta, "r");
  if (!in) {
    printf("\n Can not open file %s\n", fasta);
    exit(1);
  }

  /* FASTA pass (209-316): header lines start a contig (name = header without '>', trailing non-alphanumerics
     stripped, white space -> '_'); every alphabetic character of the other lines is a base, upper-cased. */
  size_t cap = 1 << 24, gs = 0;
  char *genome = malloc(cap);
  int n_contigs = 0, contig_cap = 64;
  int64_t *lens = malloc(sizeof(int64_t) * (size_t)contig_cap);
  char(*names)[4200] = malloc((size_t)contig_cap * 4200);
  char line[256]; /* the reference reads 255 characters at a time (fgets(sss, 256, ...)) */
  while (fgets(line, sizeof line, in)) {
    if (line[0] == '>') {
      if (n_contigs == contig_cap) {
        contig_cap *= 2;
        lens = realloc(lens, sizeof(int64_t) * (size_t)contig_cap);
        names = realloc(names, (size_t)contig_cap * 4200);
      }
      int j = (int)strlen(line);
      while (j > 1 && !isalnum((unsigned char)line[j])) line[j--] = '\0'; /* 229-233 */
      for (int i = 1; i <= j; i++) names[n_contigs][i - 1] = isspace((unsigned char)line[i]) ? '_' : line[i];
      names[n_contigs][j] = '\0';
      lens[n_contigs++] = 0;
      continue;
    }
    if (n_contigs == 0) continue; /* text before the first header */
    for (const char *p = line; *p; p++)
      if (isalpha((unsigned char)*p)) {
        if (gs + 1 >= cap) {
          cap *= 2;
          genome = realloc(genome, cap);
          if (!genome) die(" Out of memory reading the genome ");
        }
        genome[gs++] = (char)toupper((unsigned char)*p);
        lens[n_contigs - 1]++;
      }
  }
  fclose(in);
  if (n_contigs == 0) die(" No contig found in the FASTA file ");
  for (int i = 0; i < n_contigs; i++)
    if (lens[i] < 16) die(" A contig is shorter than 16 bases ");

  char path[1200];
  snprintf(path, sizeof path, "%s.seq", base);
  gzFile seq = gzopen(path, "w");
  if (!seq) die(" Could Not Open the .seq file ");
  gzbuffer(seq, 1 << 24);
  for (size_t at = 0; at < gs; at += (size_t)1 << 28) {
    const size_t n = gs - at < ((size_t)1 << 28) ? gs - at : (size_t)1 << 28;
    if (gzwrite(seq, genome + at, (unsigned)n) != (int)n) die(" Short write on the .seq file ");
  }
  gzclose(seq);

  pemap_params prm;
  pemap_default_params(&prm);
  prm.is_bisulfite = bisulfite;
  pemap_t *h = NULL;
  const int device = getenv("PEMAP_DEVICE") ? atoi(getenv("PEMAP_DEVICE")) : 0;
  /* the library refuses 2..7 contigs because MAPPING is undefined there in the reference (find_chrom quirk);
     the index itself is well defined */
  setenv("PEMAP_INDEX_ONLY", "1", 1);
  if (pemap_init_from_genome(&h, genome, lens, n_contigs, &prm, device)) {
    printf("\n pemap_init_from_genome failed: %s \n", pemap_last_error(h));
    exit(1);
  }
  uint64_t n_mers = 0;
  pemap_index_device(h, NULL, NULL, &n_mers);

  const uint64_t words = ((uint64_t)1 << 32) + 1, step = (uint64_t)1 << 26;
  uint32_t *buf = malloc(step * 4);
  if (!(getenv("PEMAP_INDEX_SKIP_IDX") && atoi(getenv("PEMAP_INDEX_SKIP_IDX")))) { /* test knob: the 16 GiB stream takes a minute to deflate */
    snprintf(path, sizeof path, "%s.idx", base); /* 334-344: exclusive prefix sums, 2^32+1 words, gz */
    gzFile idx = gzopen(path, "w");
    if (!idx) die(" Could Not Open the .idx file ");
    gzbuffer(idx, 1 << 25);
    for (uint64_t first = 0; first < words; first += step) {
      const uint64_t n = words - first < step ? words - first : step;
      if (pemap_read_pos_index(h, first, n, buf)) die(" pemap_read_pos_index failed ");
      if (gzwrite(idx, buf, (unsigned)(n * 4)) != (int)(n * 4)) die(" Short write on the .idx file ");
    }
    gzclose(idx);
  }

  snprintf(path, sizeof path, "%s.mdx", base); /* 339-340: positions grouped by k-mer, raw */
  FILE *mdx = fopen(path, "wb");
  if (!mdx) die(" Could Not Open the .mdx file ");
  for (uint64_t first = 0; first < n_mers; first += step) {
    const uint64_t n = n_mers - first < step ? n_mers - first : step;
    if (pemap_read_mers(h, first, n, buf)) die(" pemap_read_mers failed ");
    if (fwrite(buf, 4, n, mdx) != n) die(" Short write on the .mdx file ");
  }
  fclose(mdx);
  free(buf);

  snprintf(path, sizeof path, "%s.sdx", base); /* 347-351: contig_len - 15 and name per contig, then idepth */
  FILE *sdx = fopen(path, "w");
  if (!sdx) die(" Could Not Open the .sdx file ");
  fprintf(sdx, "%d\n", n_contigs);
  for (int i = 0; i < n_contigs; i++) fprintf(sdx, "%d\t%s\n", (int)(lens[i] - 15), names[i]);
  fprintf(sdx, "%d\n", 16);
  fclose(sdx);
  printf("\n Indexed %zu bases in %d contigs: %llu k-mer positions \n", gs, n_contigs, (unsigned long long)n_mers);
  pemap_destroy(h);
  return 0;
}
