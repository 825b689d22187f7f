/* gb_util.c -- small host utilities: errors, parsing, input/output streams. */
#include "gb_host.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

/* error() + exit(), Genrich.c:78-81: "Error! <msg><text>" on stderr, status 1 */
void gb_die(const char* msg, const char* suffix) {
  fprintf(stderr, "Error! %s%s\n", msg ? msg : "", suffix ? suffix : "");
  exit(EXIT_FAILURE);
}

void* gb_alloc(size_t n) {
  void* p = malloc(n ? n : 1);
  if (!p) gb_die("", "Cannot allocate memory");
  return p;
}
void* gb_realloc(void* p, size_t n) {
  void* q = realloc(p, n ? n : 1);
  if (!q) gb_die("", "Cannot allocate memory");
  return q;
}

int gb_parse_int(const char* s) {            /* getInt 117 */
  /* plain decimal numbers of up to nine digits (every numeric SAM field in practice) take a short
   * loop; anything else -- blanks, '+', overflow, garbage -- goes through strtol like the reference */
  {
    const char* p = s;
    const bool neg = *p == '-';
    if (neg) p++;
    int v = 0, nd = 0;
    while (*p >= '0' && *p <= '9' && nd < 10) { v = v * 10 + (*p - '0'); p++; nd++; }
    if (nd > 0 && nd < 10 && *p == '\0') return neg ? -v : v;
  }
  char* end;
  long v = strtol(s, &end, 10);
  if (*end != '\0') gb_die(s, ": cannot convert to int");
  return (int)v;
}
float gb_parse_float(const char* s) {        /* getFloat 106 */
  char* end;
  float v = strtof(s, &end);
  if (*end != '\0') gb_die(s, ": cannot convert to float");
  return v;
}

/* openWrite 5076-5102 */
void gb_out_open(HOut* o, const char* path, bool gz) {
  o->f = NULL;
  o->gz = NULL;
  if (path[0] == '-' && strlen(path) > 1) gb_die(path, ": output filename cannot start with '-'");
  if (gz) {
    size_t n = strlen(path);
    if ((n >= 3 && !strcmp(path + n - 3, ".gz")) || !strcmp(path, "/dev/null"))
      o->gz = gzopen(path, "w");
    else if (!strcmp(path, "-"))
      o->gz = gzdopen(fileno(stdout), "wb");
    else {
      char* p2 = (char*)gb_alloc(n + 4);
      strcpy(p2, path);
      strcat(p2, ".gz");
      o->gz = gzopen(p2, "w");
      free(p2);
    }
    if (!o->gz) gb_die(path, ": cannot open file for writing");
  } else {
    o->f = strcmp(path, "-") ? fopen(path, "w") : stdout;
    if (!o->f) gb_die(path, ": cannot open file for writing");
    if (o->f != stdout) setvbuf(o->f, NULL, _IOFBF, 1 << 20);
  }
}
void gb_out_close(HOut* o, const char* path) {
  if (o->gz && gzclose(o->gz) != Z_OK) gb_die(path, ": cannot close file");
  if (o->f && o->f != stdout && fclose(o->f)) gb_die(path, ": cannot close file");
  o->f = NULL;
  o->gz = NULL;
}
void gb_out_printf(HOut* o, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  int n = vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (n < 0) return;
  if ((size_t)n >= sizeof buf) n = sizeof buf - 1;
  if (o->gz) gzwrite(o->gz, buf, (unsigned)n);
  else fwrite(buf, 1, (size_t)n, o->f);
}

/* openRead 5132-5181 + checkBAM 5107-5126 */
bool gb_in_open(HIn* in, const char* path) {
  memset(in, 0, sizeof *in);
  bool is_stdin = !strcmp(path, "-");
  FILE* f = is_stdin ? stdin : fopen(path, "r");
  if (!f) gb_die(path, ": cannot open file for reading");
  int c0 = fgetc(f), c1 = EOF;
  if (c0 == EOF) gb_die(path, ": cannot open file for reading");
  bool gz = false;
  if ((unsigned char)c0 == 0x1F) {
    c1 = fgetc(f);
    if (c1 == EOF) gb_die(path, ": cannot open file for reading");
    gz = (unsigned char)c1 == 0x8B;
  }
  if (is_stdin) {
    if (gz) gb_die("", "Cannot pipe in gzip-compressed file (use zcat instead)");
    if (c1 != EOF) ungetc(c1, f);
    ungetc(c0, f);
    in->f = f;
    return false;
  }
  if (!gz) {
    rewind(f);
    in->f = f;
    setvbuf(f, NULL, _IOFBF, 1 << 20);
    return false;
  }
  fclose(f);
  in->gz = gzopen(path, "r");
  if (!in->gz) gb_die(path, ": cannot open file for reading");
  gzbuffer(in->gz, 1 << 20);
  in->is_gz = true;
  char magic[4] = { 'B', 'A', 'M', 1 };
  char got[4];
  int n = gzread(in->gz, got, 4);
  if (n == 4 && !memcmp(got, magic, 4)) in->is_bam = true;
  else gzrewind(in->gz);
  return true;
}
void gb_in_close(HIn* in, const char* path) {
  if (in->gz && gzclose(in->gz) != Z_OK) gb_die(path, ": cannot close file");
  if (in->f && in->f != stdin && fclose(in->f)) gb_die(path, ": cannot close file");
  in->f = NULL;
  in->gz = NULL;
}
char* gb_in_gets(HIn* in, char* line, int size) {
  return in->is_gz ? gzgets(in->gz, line, size) : fgets(line, size, in->f);
}
