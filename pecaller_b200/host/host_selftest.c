/* host_selftest.c - CPU-only checks of pemapper_gpu.c's own logic (no GPU call is made):
   * the FASTQ reader / decoder (my_gzgets semantics 2447-2483, the record scan 713-739: a '+' line, a quality line, then
     the next line starting with '@', then the sequence; quality lines that start with '@' must not be taken for headers
     out of turn; the "12 bases or fewer ends the file" rule of 663; plain and gzip input, trimming);
   * the parallel gzip writer: blocks of records deflated by several threads into concatenated gzip members must inflate
     (gzread, what pecaller does, pecaller.c:839-845) into the input bytes, across several windows.
   Built and run by tests/test_host_logic.py: gcc host_selftest.c -lpemap -lz -lpthread.  Prints "ok" lines; exit 0. */
#define main pemapper_gpu_main
#include "pemapper_gpu.c"
#undef main

static void expect(int cond, const char *what) {
  if (!cond) {
    printf("FAILED: %s\n", what);
    exit(1);
  }
}

static void write_fastq(const char *path, int gz, int n, int with_at_quality, int short_at) {
  char line[512];
  gzFile g = gz ? gzopen(path, "w") : NULL;
  FILE *f = gz ? NULL : fopen(path, "w");
  for (int i = 0; i < n; i++) {
    int len = 40 + (i * 7) % 100;
    if (i == short_at) len = 10;
    char seq[256], qual[256];
    for (int j = 0; j < len; j++) {
      seq[j] = "ACGT"[(i * 31 + j * 7 + (j * j) % 5) & 3];
      qual[j] = 'I';
    }
    if (with_at_quality && (i % 3) == 0) qual[0] = '@'; /* a legal quality character that looks like a header */
    seq[len] = qual[len] = 0;
    snprintf(line, sizeof line, "@read%d\n%s\n+\n%s\n", i, seq, qual);
    if (gz) gzputs(g, line); else fputs(line, f);
  }
  if (gz) gzclose(g); else fclose(f);
}

static long decode_all(const char *path, int trim_start, int trim_end, int check_short, char *rows, int *len, long cap) {
  decoder_t d;
  memset(&d, 0, sizeof d);
  reader_open(&d.r, path);
  d.trim_start = trim_start;
  d.trim_end = trim_end;
  d.check_short = check_short;
  decoder_prime(&d);
  long total = 0;
  for (;;) { /* small batches: the decoder must carry its pending read across calls */
    d.rows = rows + (size_t)total * ROW;
    d.len = len + total;
    d.want = 7;
    if (total + d.want > cap) d.want = cap - total;
    decoder_fill(&d);
    total += d.got;
    if (d.got < d.want || !d.pending || total >= cap) break;
  }
  reader_close(&d.r);
  return total;
}

int main(int argc, char **argv) {
  const char *dir = argc > 1 ? argv[1] : "/tmp";
  char p1[600], p2[600], p3[600];
  snprintf(p1, sizeof p1, "%s/selftest_plain.fastq", dir);
  snprintf(p2, sizeof p2, "%s/selftest.fastq.gz", dir);
  snprintf(p3, sizeof p3, "%s/selftest.pileup.gz", dir);

  /* ---- FASTQ decoding */
  const int n = 500;
  char *rows = calloc((size_t)n + 8, ROW), *rows2 = calloc((size_t)n + 8, ROW);
  int *len = calloc((size_t)n + 8, sizeof(int)), *len2 = calloc((size_t)n + 8, sizeof(int));
  write_fastq(p1, 0, n, 1, -1);
  write_fastq(p2, 1, n, 1, -1);
  long a = decode_all(p1, 0, 0, 1, rows, len, n + 8), b = decode_all(p2, 0, 0, 1, rows2, len2, n + 8);
  expect(a == n && b == n, "every record of the plain and of the gzip FASTQ is decoded");
  for (int i = 0; i < n; i++) {
    const int want = 40 + (i * 7) % 100;
    expect(len[i] == want && len2[i] == want, "read lengths");
    expect(memcmp(rows + (size_t)i * ROW, rows2 + (size_t)i * ROW, (size_t)want) == 0, "plain and gzip input give the same rows");
    for (int j = 0; j < want; j++)
      expect(rows[(size_t)i * ROW + j] == "ACGT"[(i * 31 + j * 7 + (j * j) % 5) & 3], "sequence bytes");
  }
  printf("ok fastq: %d reads, '@' quality lines not mistaken for headers, plain == gzip\n", n);
  /* the first file of a pair ends at a read of 12 bases or fewer (663) */
  write_fastq(p1, 0, n, 0, 123);
  a = decode_all(p1, 0, 0, 1, rows, len, n + 8);
  expect(a == 123, "a short read ends the first file");
  a = decode_all(p1, 0, 0, 0, rows, len, n + 8);
  expect(a == n && len[123] == 10, "the second file keeps short reads");
  /* trimming (pemapper_tsw.c: trim_from_start / trim_from_end) */
  write_fastq(p1, 0, 50, 0, -1);
  a = decode_all(p1, 5, 7, 0, rows2, len2, n + 8);
  expect(a == 50, "trimmed decode");
  for (int i = 0; i < 50; i++) {
    const int want = 40 + (i * 7) % 100;
    expect(len2[i] == want - 12, "trimmed length");
    for (int j = 0; j < len2[i]; j++)
      expect(rows2[(size_t)i * ROW + j] == "ACGT"[(i * 31 + (j + 5) * 7 + ((j + 5) * (j + 5)) % 5) & 3], "trimmed bytes");
  }
  printf("ok fastq: short-read rule and trimming\n");

  /* ---- parallel gzip members: three windows of different sizes, four threads */
  const uint64_t sizes[3] = {3 * GZ_BLOCK_RECORDS + 17, 1, 2 * GZ_BLOCK_RECORDS};
  uint64_t total = 0;
  for (int w = 0; w < 3; w++) total += sizes[w];
  pemap_record *rec = malloc(total * sizeof(pemap_record));
  uint32_t x = 12345;
  for (uint64_t i = 0; i < total; i++) {
    rec[i].pos = (uint32_t)(3 * i + 1);
    for (int k = 0; k < 6; k++) {
      x = x * 1664525u + 1013904223u;
      rec[i].c[k] = (uint16_t)((x >> 20) % 60);
    }
  }
  gzpool_t g;
  gzpool_start(&g, 4, 6);
  FILE *f = fopen(p3, "wb");
  uint64_t at = 0;
  for (int w = 0; w < 3; w++) {
    gzpool_write(&g, f, rec + at, sizes[w]);
    at += sizes[w];
  }
  fclose(f);
  gzpool_stop(&g);
  expect(g.bytes_in == total * 16 && g.bytes_out > 0 && g.bytes_out < g.bytes_in, "writer counters");
  gzFile in = gzopen(p3, "r");
  expect(in != NULL, "reopen the pileup");
  unsigned char *back = malloc(total * 16 + 16);
  uint64_t got = 0;
  for (;;) {
    int r = gzread(in, back + got, (unsigned)((total * 16 + 16 - got) > (1u << 20) ? (1u << 20) : (total * 16 + 16 - got)));
    if (r <= 0) break;
    got += (uint64_t)r;
  }
  gzclose(in);
  expect(got == total * 16, "inflated size");
  expect(memcmp(back, rec, total * 16) == 0, "inflated stream is the records, byte for byte");
  printf("ok gzip: %llu records in %d windows, %.1fx, members inflate into the record stream\n", (unsigned long long)total, 3,
         (double)g.bytes_in / (double)g.bytes_out);
  remove(p1);
  remove(p2);
  remove(p3);
  return 0;
}
