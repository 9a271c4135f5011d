/*
 * ref_wrap.c - TEST INFRASTRUCTURE ONLY.
 *
 * Builds the UNMODIFIED reference mapper into an in-process library: the reference source is
 * #included from where it lies (REF_PEMAPPER_C = /root/reference/src/pemapper.c, never copied into
 * this repo), its main() is renamed, and the few entry points below set up the reference's own
 * file-static globals from in-memory arrays (what its main() does from files at pemapper.c:411-601)
 * and then call the reference's own map_everything / initial_map / smith_waterman_align.
 * Output: oracle/_ref/libpemapper_ref.so (git-ignored; it travels to the GPU box with the snapshot).
 *
 * Used (a) to pin oracle/pemap_oracle.c stage by stage and read by read, (b) as the "reference"
 * CPU baseline of bench.py.  The product never loads it.
 */
#define main pemapper_reference_main
#include REF_PEMAPPER_C
#undef main

#include <stdint.h>

static int refw_ready = 0;
static int refw_live_threads = 0;   /* worker threads that had a batch in the last refw_map wave */

/* first-touch initialisation of the BASE_NODE array in parallel (99 GB for a 3.1 Gb genome) */
typedef struct { const char *genome; long gsize, lo, hi; } refw_init_job;
static void *refw_init_range (void *arg)
{
  refw_init_job *jb = (refw_init_job *) arg;
  long pos;
  for (pos = jb->lo; pos < jb->hi; pos++)
  {
    BASE_NODE *node = &all_base_list[pos];
    node->ref = pos < jb->gsize ? jb->genome[pos] : 'N';
    node->As = node->Cs = node->Gs = node->Ts = node->Dels = node->no_ins = 0;
    node->pos = (unsigned int) pos;
    node->ins = NULL;
  }
  return NULL;
}
static int refw_idepth = 16;
static int refw_min_dist = 0, refw_max_dist = 0;

/* Equivalent of main() 411-601 with arrays instead of files.  pos_index_ must be the dense table of
   2^32+1 entries (the inflated .idx); mers_ the .mdx; genome the inflated .seq; cstarts the n+1 prefix
   sums of the .sdx lengths (unpadded).  All borrowed for the lifetime of the process. */
int refw_init(const char *genome, long gsize, const unsigned int *cstarts, int n_contigs, unsigned int *pos_index_,
              unsigned int *mers_, double min_align, int is_bisulfite, int paired, int min_dist, int max_dist)
{
  int i, j, k;
  long pos;
  if (refw_ready)
    return -1;
  genome_size = gsize;
  max_contigs = no_contigs = n_contigs;
  contig_starts = uvector (0, max_contigs + 16);
  for (i = 0; i < max_contigs + 16; i++)
    contig_starts[i] = 0;
  for (i = 0; i <= max_contigs; i++)
    contig_starts[i] = cstarts[i];
  seq_int_vector = ivector (0, 256);
  seq_int_mat = (unsigned int ****) malloc (sizeof (unsigned int ***) * 4);
  for (i = 0; i < 4; i++)
  {
    seq_int_mat[i] = (unsigned int ***) malloc (sizeof (unsigned int **) * 4);
    for (j = 0; j < 4; j++)
    {
      seq_int_mat[i][j] = (unsigned int **) malloc (sizeof (unsigned int *) * 4);
      for (k = 0; k < 4; k++)
        seq_int_mat[i][j][k] = (unsigned int *) malloc (sizeof (unsigned int) * 4);
    }
  }
  fill_cv_mat (seq_int_vector, seq_int_mat);
  all_base_list = (BASE_NODE *) malloc ((gsize + 64) * sizeof (BASE_NODE));
  if (!all_base_list)
    return -2;
  {
    enum { NT = 16 };
    pthread_t th[NT];
    refw_init_job jobs[NT];
    long total = gsize + 64, per = (total + NT - 1) / NT;
    int nt = 0;
    for (pos = 0; pos < total; pos += per, nt++)
    {
      jobs[nt].genome = genome;
      jobs[nt].gsize = gsize;
      jobs[nt].lo = pos;
      jobs[nt].hi = pos + per < total ? pos + per : total;
      pthread_create (&th[nt], NULL, refw_init_range, &jobs[nt]);
    }
    for (i = 0; i < nt; i++)
      pthread_join (th[i], NULL);
  }
  genome_mutex_size = 1 + genome_size / genome_chunk;
  all_base_mutex = (pthread_mutex_t *) malloc (sizeof (pthread_mutex_t) * genome_mutex_size);
  for (i = 0; i < genome_mutex_size; i++)
    pthread_mutex_init (&all_base_mutex[i], NULL);
  pos_index = pos_index_;
  mers = mers_;
  /* the one-substitution table of main(), pemapper.c:546-565 */
  mismatch = imatrix (0, 255, 0, 11);
  for (i = 0; i < 256; i++)
  {
    int which = 0;
    for (j = 0; j < 4; j++)
    {
      int field = (i >> (2 * j)) & 3;
      for (k = 0; k < 4; k++)
        if (k != field)
          mismatch[i][which++] = (i & ~(3 << (2 * j))) + (k << (2 * j));
    }
  }
  MIN_ALIGN = min_align;
  IS_BISULFITE = is_bisulfite;
  pair_flag = paired;
  refw_min_dist = min_dist;
  refw_max_dist = max_dist;
  for (i = 0; i <= NEITHER_MAP; i++)
    mate_counts[i] = 0;
  gap_open = dmatrix (0, MAX_READ_LENGTH, 0, MAX_READ_LENGTH);
  gap_extend = dmatrix (0, MAX_READ_LENGTH, 0, MAX_READ_LENGTH);
  forward_bonus_matrix = dmatrix (0, MAX_READ_LENGTH, 0, MAX_READ_LENGTH);
  reverse_bonus_matrix = dmatrix (0, MAX_READ_LENGTH, 0, MAX_READ_LENGTH);
  init_bonus_matrices (match_bonus, forward_bonus_matrix, reverse_bonus_matrix, gap_open, gap_extend, MAX_READ_LENGTH);
  total_bases = total_reads = total_dist = no_dists = 0;
  refw_ready = 1;
  return 0;
}

void refw_set_params (double min_align, int paired, int min_dist, int max_dist)
{
  MIN_ALIGN = min_align;
  pair_flag = paired;
  refw_min_dist = min_dist;
  refw_max_dist = max_dist;
}

/* Run the reference's worker on n reads: batches of reads_per_thread (20,000) handed to map_everything
   on up to nthreads joinable threads (main() 616-786 uses detached threads + mutex polling). */
int refw_map (int n, const char *reads1, const int *len1, const char *reads2, const int *len2, int stride,
              unsigned int *m1, unsigned int *m2, int *mapping_type, int nthreads)
{
  int nb = (n + reads_per_thread - 1) / reads_per_thread, b, t, i;
  if (nthreads < 1)
    nthreads = 1;
  free (maps1);
  free (maps2);
  maps1 = calloc (n + 16, sizeof (unsigned int));
  maps2 = calloc (n + 16, sizeof (unsigned int));
  static PTHREAD_DATA_NODE *node_cache[256];   /* batch descriptors are reused across calls */
  static int node_cache_n = 0;
  if (nthreads > 256)
    nthreads = 256;
  PTHREAD_DATA_NODE **nodes = malloc (sizeof (PTHREAD_DATA_NODE *) * nthreads);
  pthread_t *th = malloc (sizeof (pthread_t) * nthreads);
  for (t = 0; t < nthreads; t++)
  {
    if (t >= node_cache_n)
      node_cache[node_cache_n++] = pd_node_alloc (refw_min_dist, refw_max_dist, refw_idepth, reads_per_thread);
    nodes[t] = node_cache[t];
    nodes[t]->min_dist = refw_min_dist;
    nodes[t]->max_dist = refw_max_dist;
  }
  for (b = 0; b < nb; b += nthreads)
  {
    int live = 0;
    for (t = 0; t < nthreads && b + t < nb; t++, live++)
    {
      PTHREAD_DATA_NODE *pn = nodes[t];
      int lo = (b + t) * reads_per_thread, hi = lo + reads_per_thread;
      if (hi > n)
        hi = n;
      for (i = lo; i < hi; i++)
      {
        memcpy (pn->read1[i - lo], reads1 + (size_t) i * stride, len1[i]);
        pn->read1[i - lo][len1[i]] = '\0';
        pn->len1[i - lo] = len1[i];
        if (pair_flag)
        {
          memcpy (pn->read2[i - lo], reads2 + (size_t) i * stride, len2[i]);
          pn->read2[i - lo][len2[i]] = '\0';
          pn->len2[i - lo] = len2[i];
        }
        pn->read_no[i - lo] = i;
        pn->m1[i - lo] = pn->m2[i - lo] = 0;   /* reused descriptor: as freshly allocated */
        pn->mapping_type[i - lo] = 0;
      }
      pn->this_tot = hi - lo;
      pn->tid = t;
      pthread_mutex_lock (&pn->mutex);
      if (pthread_create (&th[t], NULL, map_everything, (void *) pn))
        return -1;
    }
    if (live > refw_live_threads || b == 0)
      refw_live_threads = live;
    for (t = 0; t < live; t++)
    {
      PTHREAD_DATA_NODE *pn = nodes[t];
      int lo = (b + t) * reads_per_thread, hi = lo + reads_per_thread;
      if (hi > n)
        hi = n;
      pthread_join (th[t], NULL);
      for (i = lo; i < hi; i++)
      {
        m1[i] = pn->m1[i - lo];
        m2[i] = pn->m2[i - lo];
        mapping_type[i] = pn->mapping_type[i - lo];
      }
    }
  }
  free (nodes);
  free (th);
  return 0;
}

int refw_last_live_threads (void) { return refw_live_threads; }
long refw_mate_count (int type) { return mate_counts[type]; }
long refw_total_reads (void) { return total_reads; }
long refw_total_bases (void) { return total_bases; }
long refw_total_dist (void) { return total_dist; }
long refw_no_dists (void) { return no_dists; }

/* the writer loop of main(), pemapper.c:828-842, into memory: 16-byte records */
unsigned long refw_records (unsigned char *out, unsigned long cap)
{
  unsigned long n = 0;
  unsigned int pos;
  for (pos = 0; pos < genome_size; pos++)
  {
    BASE_NODE *b = &all_base_list[pos];
    if (b->As + b->Cs + b->Gs + b->Ts + b->Dels + b->no_ins > 0)
    {
      if (n < cap)
      {
        unsigned short rc[6] = { b->As, b->Cs, b->Gs, b->Ts, b->Dels, b->no_ins };
        memcpy (out + 16 * n, &pos, 4);
        memcpy (out + 16 * n + 4, rc, 12);
      }
      n++;
    }
  }
  return n;
}

unsigned long refw_n_insertions (void)
{
  unsigned long n = 0;
  long pos;
  for (pos = 0; pos < genome_size; pos++)
    n += all_base_list[pos].no_ins;
  return n;
}

/* insertion strings in site order; each entry: pos, string copied into buf rows of width w */
unsigned long refw_insertions (unsigned int *pos_out, char *buf, int w, unsigned long cap)
{
  unsigned long n = 0;
  long pos;
  int j;
  for (pos = 0; pos < genome_size; pos++)
    for (j = 0; j < all_base_list[pos].no_ins; j++)
    {
      if (n < cap)
      {
        pos_out[n] = (unsigned int) pos;
        strncpy (buf + (size_t) n * w, all_base_list[pos].ins[j], w - 1);
        buf[(size_t) n * w + w - 1] = '\0';
      }
      n++;
    }
  return n;
}

void refw_reset_counts (void)
{
  long pos;
  int j;
  for (pos = 0; pos < genome_size; pos++)
  {
    BASE_NODE *b = &all_base_list[pos];
    for (j = 0; j < b->no_ins; j++)
      free (b->ins[j]);
    if (b->no_ins)
      free (b->ins);
    b->ins = NULL;
    b->As = b->Cs = b->Gs = b->Ts = b->Dels = b->no_ins = 0;
  }
  for (j = 0; j <= NEITHER_MAP; j++)
    mate_counts[j] = 0;
  total_bases = total_reads = total_dist = no_dists = 0;
}

/* stage probes: the reference's own initial_map and smith_waterman_align */
int refw_initial_map (const char *read, int len, unsigned int *spots, char *orients)
{
  char **ir = cmatrix (0, 1, 0, MAX_READ_LENGTH);
  int hits;
  memcpy (ir[0], read, len);
  ir[0][len] = '\0';
  reverse_transcribe (ir[0], ir[1], len);
  hits = initial_map (ir, len, orients, spots, refw_idepth);
  free_cmatrix (ir, 0, 1, 0, MAX_READ_LENGTH);
  return hits;
}

double refw_sw_align (unsigned int win_start, int blen, const char *seq, int mm, int *start3)
{
  int ms = MAX_READ_LENGTH, i, j;
  double ***S = (double ***) malloc (sizeof (double **) * 3), score;
  double gopen_penalty = 2.0 * match_bonus, ge = match_bonus / 36.0;
  for (i = 0; i < 3; i++)
  {
    S[i] = (double **) malloc (sizeof (double *) * ms);
    for (j = 0; j < ms; j++)
      S[i][j] = (double *) malloc (sizeof (double) * ms);
  }
  /* borders as init_penalty_matrices (2051-2095) lays them out for one candidate slot */
  S[0][0][0] = 0.0;
  S[1][0][0] = 0.0;
  S[2][0][0] = -1.0 * gopen_penalty;
  for (i = 1; i < ms; i++)
  {
    S[0][0][i] = S[1][0][i] = S[2][0][i] = -(gopen_penalty + (double) (i - 1) * ge);
    S[0][i][0] = S[0][0][0];
    S[1][i][0] = S[1][0][0];
    S[2][i][0] = S[2][0][0];
  }
  score = smith_waterman_align (&all_base_list[win_start], blen, (char *) seq, mm, S, forward_bonus_matrix, gap_open,
                                gap_extend, start3);
  for (i = 0; i < 3; i++)
  {
    for (j = 0; j < ms; j++)
      free (S[i][j]);
    free (S[i]);
  }
  free (S);
  return score;
}
