/* gtcompat.c -- minimal stand-in for the parts of the GenomeTools C library
   that the gt Scaffolder sources use (containers, strings, error object,
   output file, FASTA reader, logger, N-statistics).

   GenomeTools is an external dependency of the reference that is neither
   vendored in it nor installed here (reference README.md:30-33,
   src/Makefile:6-9).  None of the hot-path arithmetic lives in it; the only
   result-affecting behaviours are gt_str_cmp ordering (strcmp), FASTA
   header/length extraction and printf formatting.  Everything below is
   written from the public GenomeTools API descriptions, not copied.

   Used by: (1) the oracle build, which compiles the reference sources
   unmodified against these headers; (2) the drop-in boundary, which allocates
   GtScaffolderGraph members with gt_malloc/GtStr so that the reference's
   gt_scaffolder_graph_delete can free them.  In a real deployment both link
   libgenometools instead. */
#include <ctype.h>
#include <stdarg.h>
#include <string.h>

#include "core/array_api.h"
#include "core/cstr_api.h"
#include "core/error.h"
#include "core/fasta_reader_rec.h"
#include "core/file_api.h"
#include "core/init_api.h"
#include "core/logger.h"
#include "core/ma_api.h"
#include "core/queue_api.h"
#include "core/str_api.h"
#include "extended/assembly_stats_calculator.h"

/* ---------------------------------------------------------------- memory */

static void gtcompat_oom(size_t size, const char *file, int line)
{
  fprintf(stderr, "gtcompat: cannot allocate %zu bytes (%s:%d)\n", size, file,
          line);
  exit(EXIT_FAILURE);
}

void *gt_malloc_mem(size_t size, const char *file, int line)
{
  void *p = malloc(size ? size : 1);
  if (p == NULL) gtcompat_oom(size, file, line);
  return p;
}

void *gt_calloc_mem(size_t nmemb, size_t size, const char *file, int line)
{
  void *p = calloc(nmemb ? nmemb : 1, size ? size : 1);
  if (p == NULL) gtcompat_oom(nmemb * size, file, line);
  return p;
}

void *gt_realloc_mem(void *ptr, size_t size, const char *file, int line)
{
  void *p = realloc(ptr, size ? size : 1);
  if (p == NULL) gtcompat_oom(size, file, line);
  return p;
}

void gt_free_mem(void *ptr) { free(ptr); }

char *gt_cstr_dup(const char *cstr)
{
  size_t n = strlen(cstr) + 1;
  char *copy = gt_malloc(n);
  memcpy(copy, cstr, n);
  return copy;
}

/* ----------------------------------------------------------------- error */

struct GtError {
  char msg[1024];
  bool is_set;
};

GtError *gt_error_new(void)
{
  GtError *err = gt_calloc(1, sizeof (*err));
  return err;
}

void gt_error_set(GtError *err, const char *format, ...)
{
  va_list ap;
  if (err == NULL) return;
  va_start(ap, format);
  vsnprintf(err->msg, sizeof (err->msg), format, ap);
  va_end(ap);
  err->is_set = true;
}

bool gt_error_is_set(const GtError *err) { return err != NULL && err->is_set; }

void gt_error_unset(GtError *err)
{
  if (err != NULL) {
    err->is_set = false;
    err->msg[0] = '\0';
  }
}

const char *gt_error_get(const GtError *err)
{
  gt_assert(err != NULL);
  return err->msg;
}

void gt_error_delete(GtError *err) { gt_free(err); }

/* ---------------------------------------------------------------- string */

struct GtStr {
  char *cstr;
  GtUword length, allocated;
};

static void gt_str_reserve(GtStr *s, GtUword need)
{
  if (need + 1 > s->allocated) {
    GtUword cap = s->allocated ? s->allocated : 32;
    while (cap < need + 1) cap *= 2;
    s->cstr = gt_realloc(s->cstr, cap);
    s->allocated = cap;
  }
}

GtStr *gt_str_new(void)
{
  GtStr *s = gt_malloc(sizeof (*s));
  s->cstr = NULL;
  s->length = s->allocated = 0;
  gt_str_reserve(s, 0);
  s->cstr[0] = '\0';
  return s;
}

GtStr *gt_str_new_cstr(const char *cstr)
{
  GtStr *s = gt_str_new();
  if (cstr != NULL) gt_str_set(s, cstr);
  return s;
}

GtStr *gt_str_clone(const GtStr *str)
{
  gt_assert(str != NULL);
  return gt_str_new_cstr(str->cstr);
}

void gt_str_set(GtStr *s, const char *cstr)
{
  size_t n;
  gt_assert(s != NULL);
  if (cstr == NULL) cstr = "";
  n = strlen(cstr);
  gt_str_reserve(s, n);
  memcpy(s->cstr, cstr, n + 1);
  s->length = n;
}

void gt_str_append_cstr(GtStr *s, const char *cstr)
{
  size_t n = strlen(cstr);
  gt_str_reserve(s, s->length + n);
  memcpy(s->cstr + s->length, cstr, n + 1);
  s->length += n;
}

char *gt_str_get(const GtStr *s)
{
  gt_assert(s != NULL);
  return s->cstr;
}

GtUword gt_str_length(const GtStr *s) { return s ? s->length : 0; }

int gt_str_cmp(const GtStr *a, const GtStr *b)
{
  gt_assert(a != NULL && b != NULL);
  if (a == b) return 0;
  return strcmp(a->cstr, b->cstr);
}

void gt_str_delete(GtStr *s)
{
  if (s == NULL) return;
  gt_free(s->cstr);
  gt_free(s);
}

/* ----------------------------------------------------------------- array */

struct GtArray {
  char *space;
  GtUword next_free, allocated;
  size_t size_of_elem;
};

GtArray *gt_array_new(size_t size_of_elem)
{
  GtArray *a = gt_calloc(1, sizeof (*a));
  gt_assert(size_of_elem > 0);
  a->size_of_elem = size_of_elem;
  return a;
}

void gt_array_add_elem(GtArray *a, void *elem, size_t size_of_elem)
{
  gt_assert(a != NULL && elem != NULL);
  gt_assert(a->size_of_elem == size_of_elem);
  if (a->next_free == a->allocated) {
    a->allocated = a->allocated ? 2 * a->allocated : 16;
    a->space = gt_realloc(a->space, a->allocated * a->size_of_elem);
  }
  memcpy(a->space + a->next_free * a->size_of_elem, elem, a->size_of_elem);
  a->next_free++;
}

void *gt_array_get(const GtArray *a, GtUword idx)
{
  gt_assert(a != NULL && idx < a->next_free);
  return a->space + idx * a->size_of_elem;
}

void *gt_array_pop(GtArray *a)
{
  gt_assert(a != NULL && a->next_free > 0);
  a->next_free--;
  return a->space + a->next_free * a->size_of_elem;
}

GtUword gt_array_size(const GtArray *a) { return a ? a->next_free : 0; }

void gt_array_reset(GtArray *a)
{
  gt_assert(a != NULL);
  a->next_free = 0;
}

void gt_array_delete(GtArray *a)
{
  if (a == NULL) return;
  gt_free(a->space);
  gt_free(a);
}

/* ----------------------------------------------------------------- queue */

struct GtQueue {
  void **ring;
  GtUword head, count, allocated;
};

GtQueue *gt_queue_new(void) { return gt_calloc(1, sizeof (GtQueue)); }

void gt_queue_add(GtQueue *q, void *elem)
{
  gt_assert(q != NULL);
  if (q->count == q->allocated) {
    GtUword i, cap = q->allocated ? 2 * q->allocated : 64;
    void **ring = gt_malloc(cap * sizeof (*ring));
    for (i = 0; i < q->count; i++)
      ring[i] = q->ring[(q->head + i) % q->allocated];
    gt_free(q->ring);
    q->ring = ring;
    q->head = 0;
    q->allocated = cap;
  }
  q->ring[(q->head + q->count) % q->allocated] = elem;
  q->count++;
}

void *gt_queue_get(GtQueue *q)
{
  void *elem;
  gt_assert(q != NULL && q->count > 0);
  elem = q->ring[q->head];
  q->head = (q->head + 1) % q->allocated;
  q->count--;
  return elem;
}

GtUword gt_queue_size(const GtQueue *q)
{
  gt_assert(q != NULL);
  return q->count;
}

void gt_queue_delete(GtQueue *q)
{
  if (q == NULL) return;
  gt_free(q->ring);
  gt_free(q);
}

/* ------------------------------------------------------------------ file */

struct GtFile {
  FILE *fp;
};

GtFile *gt_file_new(const char *path, const char *mode, GtError *err)
{
  GtFile *f;
  FILE *fp = fopen(path, mode);
  if (fp == NULL) {
    gt_error_set(err, "cannot open file '%s'", path);
    return NULL;
  }
  f = gt_malloc(sizeof (*f));
  f->fp = fp;
  return f;
}

void gt_file_xprintf(GtFile *file, const char *format, ...)
{
  va_list ap;
  FILE *fp = file ? file->fp : stdout;
  va_start(ap, format);
  if (vfprintf(fp, format, ap) < 0) {
    fprintf(stderr, "gtcompat: write error\n");
    exit(EXIT_FAILURE);
  }
  va_end(ap);
}

void gt_file_delete(GtFile *file)
{
  if (file == NULL) return;
  fclose(file->fp);
  gt_free(file);
}

/* ---------------------------------------------------------- FASTA reader */

struct GtFastaReader {
  GtStr *filename;
};

GtFastaReader *gt_fasta_reader_rec_new(GtStr *sequence_filename)
{
  GtFastaReader *r = gt_malloc(sizeof (*r));
  r->filename = gt_str_clone(sequence_filename);
  return r;
}

int gt_fasta_reader_run(GtFastaReader *reader,
                        GtFastaReaderProcDescription proc_description,
                        GtFastaReaderProcSequencePart proc_sequence_part,
                        GtFastaReaderProcSequenceLength proc_sequence_length,
                        void *data, GtError *err)
{
  FILE *fp;
  GtStr *desc;
  char *line = NULL;
  size_t cap = 0;
  ssize_t n;
  GtUword seqlen = 0;
  bool in_entry = false;
  int had_err = 0;

  gt_assert(reader != NULL);
  fp = fopen(gt_str_get(reader->filename), "rb");
  if (fp == NULL) {
    gt_error_set(err, "cannot open file '%s'", gt_str_get(reader->filename));
    return -1;
  }
  desc = gt_str_new();
  while (!had_err && (n = getline(&line, &cap, fp)) != -1) {
    while (n > 0 && (line[n-1] == '\n' || line[n-1] == '\r')) line[--n] = '\0';
    if (line[0] == '>') {
      if (in_entry && proc_sequence_length != NULL)
        had_err = proc_sequence_length(seqlen, data, err);
      in_entry = true;
      seqlen = 0;
      gt_str_set(desc, line + 1);
      if (!had_err && proc_description != NULL)
        had_err = proc_description(gt_str_get(desc), gt_str_length(desc), data,
                                   err);
    }
    else if (in_entry) {
      GtUword i, kept = 0;
      for (i = 0; i < (GtUword) n; i++)
        if (!isspace((unsigned char) line[i])) line[kept++] = line[i];
      line[kept] = '\0';
      if (kept > 0 && proc_sequence_part != NULL)
        had_err = proc_sequence_part(line, kept, data, err);
      seqlen += kept;
    }
    else if (n > 0) {
      gt_error_set(err, "file '%s' does not start with '>'",
                   gt_str_get(reader->filename));
      had_err = -1;
    }
  }
  if (!had_err && in_entry && proc_sequence_length != NULL)
    had_err = proc_sequence_length(seqlen, data, err);
  free(line);
  gt_str_delete(desc);
  fclose(fp);
  return had_err;
}

void gt_fasta_reader_delete(GtFastaReader *reader)
{
  if (reader == NULL) return;
  gt_str_delete(reader->filename);
  gt_free(reader);
}

/* ------------------------------------------------------- init and logger */

void gt_lib_init(void) {}
int gt_lib_clean(void) { return 0; }

struct GtLogger {
  bool enabled;
  char prefix[64];
  FILE *target;
};

GtLogger *gt_logger_new(bool enabled, const char *prefix, FILE *target)
{
  GtLogger *l = gt_calloc(1, sizeof (*l));
  l->enabled = enabled;
  snprintf(l->prefix, sizeof (l->prefix), "%s", prefix ? prefix : "");
  l->target = target;
  return l;
}

void gt_logger_log(GtLogger *logger, const char *format, ...)
{
  va_list ap;
  if (logger == NULL || !logger->enabled) return;
  fputs(logger->prefix, logger->target);
  va_start(ap, format);
  vfprintf(logger->target, format, ap);
  va_end(ap);
  fputc('\n', logger->target);
}

void gt_logger_delete(GtLogger *logger) { gt_free(logger); }

/* ---------------------------------------------------- assembly statistics */

struct GtAssemblyStatsCalculator {
  GtUword *lengths, count, allocated, total, nstat;
};

GtAssemblyStatsCalculator *gt_assembly_stats_calculator_new(void)
{
  GtAssemblyStatsCalculator *c = gt_calloc(1, sizeof (*c));
  c->nstat = 50;
  return c;
}

void gt_assembly_stats_calculator_add(GtAssemblyStatsCalculator *c,
                                      GtUword length)
{
  if (c->count == c->allocated) {
    c->allocated = c->allocated ? 2 * c->allocated : 64;
    c->lengths = gt_realloc(c->lengths, c->allocated * sizeof (GtUword));
  }
  c->lengths[c->count++] = length;
  c->total += length;
}

void gt_assembly_stats_calculator_nstat(GtAssemblyStatsCalculator *c,
                                        GtUword n)
{
  c->nstat = n;
}

static int gtcompat_cmp_desc(const void *a, const void *b)
{
  GtUword x = *(const GtUword *) a, y = *(const GtUword *) b;
  return x < y ? 1 : (x > y ? -1 : 0);
}

void gt_assembly_stats_calculator_show(GtAssemblyStatsCalculator *c,
                                       GtLogger *logger)
{
  GtUword i, acc = 0, nval = 0;
  if (c->count > 0) {
    qsort(c->lengths, c->count, sizeof (GtUword), gtcompat_cmp_desc);
    for (i = 0; i < c->count; i++) {
      acc += c->lengths[i];
      if (acc * 100 >= c->total * c->nstat) {
        nval = c->lengths[i];
        break;
      }
    }
  }
  gt_logger_log(logger, "number of scaffolds: " GT_WU, c->count);
  gt_logger_log(logger, "total length: " GT_WU, c->total);
  gt_logger_log(logger, "N" GT_WU ": " GT_WU, c->nstat, nval);
}

void gt_assembly_stats_calculator_delete(GtAssemblyStatsCalculator *c)
{
  if (c == NULL) return;
  gt_free(c->lengths);
  gt_free(c);
}
