/*
 * fake_asterisk.c -- TEST INFRASTRUCTURE: the slice of the Asterisk runtime that the reference's unchanged
 * Asterisk-facing files call (see include/asterisk.h), plus the harness entry points (fake_*) through which
 * tests/test_gpu_asterisk_dropin.py plays the roles of the module loader, the config directory, a PBX thread
 * with a channel, and the CLI.  ast_json_* is a thin wrapper over the real libjansson.so.4, as in Asterisk.
 */
#include <asterisk.h>
#include <asterisk/app.h>
#include <asterisk/channel.h>
#include <asterisk/cli.h>
#include <asterisk/config.h>
#include <asterisk/file.h>
#include <asterisk/json.h>
#include <asterisk/logger.h>
#include <asterisk/module.h>
#include <asterisk/pbx.h>
#include <jansson.h>

#undef opendir
#undef mkdir
#include <dirent.h>
#include <pthread.h>
#include <sys/stat.h>

/* ------------------------------------------------------------------------------------------ root remap */
static const char *fake_root(void) {
  const char *r = getenv("FAKE_AST_ROOT");
  return (r && *r) ? r : "/tmp/fake_asterisk";
}
static void remap(const char *path, char *out, size_t cap) {
  if (strncmp(path, "/var/lib/asterisk", 17) == 0 || strncmp(path, "/etc/asterisk", 13) == 0) snprintf(out, cap, "%s%s", fake_root(), path);
  else snprintf(out, cap, "%s", path);
}
DIR *fake_ast_opendir(const char *name) {
  char p[4096];
  remap(name, p, sizeof p);
  return opendir(p);
}
int fake_ast_mkdir(const char *path, mode_t mode) {
  char p[4096];
  remap(path, p, sizeof p);
  return mkdir(p, mode);
}

/* ------------------------------------------------------------------------------------------ logging */
static pthread_mutex_t g_log_mu = PTHREAD_MUTEX_INITIALIZER;
static char *g_log_buf;
static size_t g_log_len, g_log_cap;
static int g_log_stderr_level = LOG_WARNING;
void fake_ast_log(int level, const char *file, int line, const char *func, const char *fmt, ...) {
  static const char *names[] = {"DEBUG", "?", "NOTICE", "WARNING", "ERROR", "VERBOSE"};
  char msg[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof msg, fmt, ap);
  va_end(ap);
  char linebuf[2300];
  const char *base = strrchr(file, '/');
  const int n = snprintf(linebuf, sizeof linebuf, "[%s] %s:%d %s: %s", names[level >= 0 && level <= 5 ? level : 1], base ? base + 1 : file, line, func, msg);
  pthread_mutex_lock(&g_log_mu);
  if (g_log_len + (size_t)n + 1 > g_log_cap) {
    g_log_cap = (g_log_len + (size_t)n + 1) * 2 + 4096;
    g_log_buf = realloc(g_log_buf, g_log_cap);
  }
  memcpy(g_log_buf + g_log_len, linebuf, (size_t)n + 1);
  g_log_len += (size_t)n;
  pthread_mutex_unlock(&g_log_mu);
  if ((level == LOG_WARNING || level == LOG_ERROR) && g_log_stderr_level <= level) fputs(linebuf, stderr);
}
/* harness: the log collected so far (and reset) */
char *fake_take_log(void) {
  pthread_mutex_lock(&g_log_mu);
  char *r = g_log_buf ? g_log_buf : strdup("");
  g_log_buf = NULL, g_log_len = g_log_cap = 0;
  pthread_mutex_unlock(&g_log_mu);
  return r;
}
void fake_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------ ast_json over jansson */
#define J(x) ((json_t *)(x))
#define A(x) ((struct ast_json *)(x))
struct ast_json *ast_json_ref(struct ast_json *v) {
  if (v && J(v)->refcount != (size_t)-1) __atomic_add_fetch(&J(v)->refcount, 1, __ATOMIC_ACQUIRE);
  return v;
}
void ast_json_unref(struct ast_json *v) {
  if (v && J(v)->refcount != (size_t)-1 && __atomic_sub_fetch(&J(v)->refcount, 1, __ATOMIC_RELEASE) == 0) json_delete(J(v));
}
enum ast_json_type ast_json_typeof(const struct ast_json *v) { return (enum ast_json_type)J(v)->type; }
struct ast_json *ast_json_null(void) { return A(json_null()); }
struct ast_json *ast_json_string_create(const char *s) { return A(json_string(s)); }
const char *ast_json_string_get(const struct ast_json *s) { return (s && J(s)->type == JSON_STRING) ? json_string_value(J(s)) : NULL; }
struct ast_json *ast_json_integer_create(intmax_t v) { return A(json_integer((json_int_t)v)); }
intmax_t ast_json_integer_get(const struct ast_json *v) { return (v && J(v)->type == JSON_INTEGER) ? (intmax_t)json_integer_value(J(v)) : 0; }
struct ast_json *ast_json_real_create(double v) { return A(json_real(v)); }
double ast_json_real_get(const struct ast_json *v) { return (v && J(v)->type == JSON_REAL) ? json_real_value(J(v)) : 0.0; }
struct ast_json *ast_json_array_create(void) { return A(json_array()); }
size_t ast_json_array_size(const struct ast_json *a) { return (a && J(a)->type == JSON_ARRAY) ? json_array_size(J(a)) : 0; }
struct ast_json *ast_json_array_get(const struct ast_json *a, size_t i) { return (a && J(a)->type == JSON_ARRAY) ? A(json_array_get(J(a), i)) : NULL; }
int ast_json_array_append(struct ast_json *a, struct ast_json *v) { return json_array_append_new(J(a), J(v)); }
int ast_json_array_remove(struct ast_json *a, size_t i) { return json_array_remove(J(a), i); }
struct ast_json *ast_json_object_create(void) { return A(json_object()); }
struct ast_json *ast_json_object_get(struct ast_json *o, const char *k) { return (o && k && J(o)->type == JSON_OBJECT) ? A(json_object_get(J(o), k)) : NULL; }
int ast_json_object_set(struct ast_json *o, const char *k, struct ast_json *v) { return json_object_set_new(J(o), k, J(v)); }
struct ast_json_iter *ast_json_object_iter(struct ast_json *o) { return (struct ast_json_iter *)json_object_iter(J(o)); }
struct ast_json_iter *ast_json_object_iter_next(struct ast_json *o, struct ast_json_iter *it) { return (struct ast_json_iter *)json_object_iter_next(J(o), it); }
const char *ast_json_object_iter_key(struct ast_json_iter *it) { return json_object_iter_key(it); }
struct ast_json *ast_json_object_iter_value(struct ast_json_iter *it) { return A(json_object_iter_value(it)); }
struct ast_json *ast_json_pack(char const *format, ...) {
  va_list ap;
  va_start(ap, format);
  json_error_t err;
  json_t *r = json_vpack_ex(&err, 0, format, ap);
  va_end(ap);
  return A(r);
}
struct ast_json *ast_json_deep_copy(const struct ast_json *v) { return A(json_deep_copy(J(v))); }
struct ast_json *ast_json_load_string(const char *input, struct ast_json_error *error) {
  (void)error;
  json_error_t err;
  return A(json_loads(input, JSON_DECODE_ANY, &err)); /* Asterisk builds jansson calls with JSON_DECODE_ANY semantics for scalars */
}
char *ast_json_dump_string_format(struct ast_json *root, enum ast_json_encoding_format format) {
  return json_dumps(J(root), (format == AST_JSON_COMPACT ? JSON_COMPACT : JSON_INDENT(2)) | JSON_ENCODE_ANY);
}

/* ------------------------------------------------------------------------------------------ config files */
struct fake_cat {
  char *name;
  struct ast_variable *vars, *last;
  struct fake_cat *next;
};
struct ast_config {
  struct fake_cat *cats, *last;
};
static char *trim(char *s) {
  while (*s == ' ' || *s == '\t') s++;
  char *e = s + strlen(s);
  while (e > s && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\n' || e[-1] == '\r')) *--e = '\0';
  return s;
}
struct ast_config *fake_ast_config_load(const char *filename, struct ast_flags flags) {
  (void)flags;
  char path[4096];
  snprintf(path, sizeof path, "%s/etc/asterisk/%s", fake_root(), filename);
  FILE *f = fopen(path, "r");
  if (!f) return CONFIG_STATUS_FILEMISSING;
  struct ast_config *cfg = calloc(1, sizeof *cfg);
  char line[2048];
  while (fgets(line, sizeof line, f)) {
    char *c = strchr(line, ';');
    if (c) *c = '\0';
    char *s = trim(line);
    if (!*s) continue;
    if (*s == '[') {
      char *e = strchr(s, ']');
      if (!e) continue;
      *e = '\0';
      struct fake_cat *cat = calloc(1, sizeof *cat);
      cat->name = strdup(trim(s + 1));
      if (cfg->last) cfg->last->next = cat; else cfg->cats = cat;
      cfg->last = cat;
    } else if (cfg->last) {
      char *eq = strchr(s, '=');
      if (!eq) continue;
      *eq = '\0';
      char *val = eq + 1;
      if (*val == '>') val++;
      struct ast_variable *v = calloc(1, sizeof *v);
      v->name = strdup(trim(s)), v->value = strdup(trim(val));
      if (cfg->last->last) cfg->last->last->next = v; else cfg->last->vars = v;
      cfg->last->last = v;
    }
  }
  fclose(f);
  return cfg;
}
void ast_config_destroy(struct ast_config *cfg) {
  if (!cfg || cfg == CONFIG_STATUS_FILEINVALID || cfg == CONFIG_STATUS_FILEUNCHANGED) return;
  for (struct fake_cat *c = cfg->cats; c;) {
    for (struct ast_variable *v = c->vars; v;) {
      struct ast_variable *n = v->next;
      free((char *)v->name), free((char *)v->value), free(v);
      v = n;
    }
    struct fake_cat *n = c->next;
    free(c->name), free(c);
    c = n;
  }
  free(cfg);
}
char *ast_category_browse(struct ast_config *cfg, const char *prev) {
  if (!cfg) return NULL;
  if (!prev) return cfg->cats ? cfg->cats->name : NULL;
  for (struct fake_cat *c = cfg->cats; c; c = c->next)
    if (c->name == prev || strcmp(c->name, prev) == 0) return c->next ? c->next->name : NULL;
  return NULL;
}
struct ast_variable *ast_variable_browse(const struct ast_config *cfg, const char *category) {
  for (struct fake_cat *c = cfg ? cfg->cats : NULL; c; c = c->next)
    if (strcmp(c->name, category) == 0) return c->vars;
  return NULL;
}

/* ------------------------------------------------------------------------------------------ CLI */
static struct ast_cli_entry *g_cli[64];
static int g_ncli;
void ast_cli(int fd, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vdprintf(fd, fmt, ap);
  va_end(ap);
}
int ast_cli_register_multiple(struct ast_cli_entry *e, int len) {
  for (int i = 0; i < len; i++) {
    e[i].handler(&e[i], CLI_INIT, NULL); /* fills command / usage */
    if (g_ncli < 64) g_cli[g_ncli++] = &e[i];
  }
  return 0;
}
int ast_cli_unregister_multiple(struct ast_cli_entry *e, int len) {
  for (int i = 0; i < len; i++)
    for (int k = 0; k < g_ncli; k++)
      if (g_cli[k] == &e[i]) g_cli[k] = g_cli[--g_ncli], k--;
  return 0;
}
/* harness: run one CLI line; output goes to fd; returns 0 success, 1 usage shown, 2 failure, -1 no such command */
int fake_cli_run(const char *line, int fd) {
  char *copy = strdup(line), *argv[32], *save = NULL;
  int argc = 0;
  for (char *t = strtok_r(copy, " \t", &save); t && argc < 31; t = strtok_r(NULL, " \t", &save)) argv[argc++] = t;
  argv[argc] = NULL;
  int rc = -1;
  for (int k = 0; k < g_ncli && rc < 0; k++) {
    char *cmd = strdup(g_cli[k]->command ? g_cli[k]->command : ""), *s2 = NULL;
    int n = 0, ok = 1;
    for (char *t = strtok_r(cmd, " ", &s2); t; t = strtok_r(NULL, " ", &s2), n++)
      if (n >= argc || strcmp(argv[n], t) != 0) ok = 0;
    free(cmd);
    if (!ok || n == 0) continue;
    struct ast_cli_args a = {.fd = fd, .argc = argc, .argv = (const char *const *)argv, .line = line};
    char *r = g_cli[k]->handler(g_cli[k], CLI_HANDLER, &a);
    if (r == CLI_SHOWUSAGE) ast_cli(fd, "%s", g_cli[k]->usage ? g_cli[k]->usage : "");
    rc = r == CLI_SUCCESS ? 0 : r == CLI_SHOWUSAGE ? 1 : 2;
  }
  free(copy);
  return rc;
}
int fake_cli_count(void) { return g_ncli; }

/* ------------------------------------------------------------------------------------------ applications, channels */
struct fake_app {
  char name[64];
  int (*exec)(struct ast_channel *, const char *);
};
static struct fake_app g_apps[16];
static int g_napps;
int ast_register_application2(const char *app, int (*execute)(struct ast_channel *, const char *), const char *synopsis, const char *description, void *mod) {
  (void)synopsis, (void)description, (void)mod;
  if (g_napps >= 16) return -1;
  snprintf(g_apps[g_napps].name, sizeof g_apps[g_napps].name, "%s", app);
  g_apps[g_napps++].exec = execute;
  return 0;
}
int ast_unregister_application(const char *app) {
  for (int i = 0; i < g_napps; i++)
    if (strcmp(g_apps[i].name, app) == 0) { g_apps[i] = g_apps[--g_napps]; return 0; }
  return -1;
}
unsigned int fake_ast_separate_args(char *buf, char delim, char **array, int arraylen) {
  unsigned int argc = 0;
  if (!buf || !array || arraylen <= 0) return 0;
  memset(array, 0, sizeof(char *) * (size_t)arraylen);
  char *p = buf;
  while (argc < (unsigned)arraylen) {
    array[argc++] = p;
    char *d = (argc < (unsigned)arraylen) ? strchr(p, delim) : NULL; /* the last slot keeps the remainder */
    if (!d) break;
    *d = '\0';
    p = d + 1;
  }
  return argc;
}

/* A scripted channel: PCM16 audio delivered as 20 ms voice frames on a VIRTUAL clock (every ast_read advances it
 * by the frame's duration), optional non-voice frames and a hangup position. */
struct ast_channel {
  enum ast_channel_state state;
  int answered;
  const int16_t *pcm;
  long n_samples, pos;
  int rate, samples_per_frame;
  long hangup_at;     /* sample position at which ast_read returns NULL (< 0: never) */
  int control_every;  /* every n-th frame is preceded by a non-voice frame (0: none) */
  long frames_read;
  long clock_ms;
  char vars[32][2][512];
  int nvars;
};
static __thread struct ast_channel *t_chan; /* the channel of the PBX thread (for the virtual clock) */

struct ast_channel *fake_channel_new(const int16_t *pcm, long n_samples, int rate, int state_up, long hangup_at, int control_every) {
  struct ast_channel *c = calloc(1, sizeof *c);
  c->state = state_up ? AST_STATE_UP : AST_STATE_RING;
  c->pcm = pcm, c->n_samples = n_samples, c->rate = rate, c->samples_per_frame = rate / 50;
  c->hangup_at = hangup_at, c->control_every = control_every;
  return c;
}
void fake_channel_free(struct ast_channel *c) { free(c); }
const char *fake_channel_var(struct ast_channel *c, const char *name) {
  for (int i = c->nvars - 1; i >= 0; i--)
    if (strcmp(c->vars[i][0], name) == 0) return c->vars[i][1];
  return NULL;
}
int fake_channel_answered(struct ast_channel *c) { return c->answered; }
long fake_channel_samples_read(struct ast_channel *c) { return c->pos; }
int fake_app_exec(const char *app, struct ast_channel *chan, const char *data) {
  for (int i = 0; i < g_napps; i++)
    if (strcmp(g_apps[i].name, app) == 0) {
      t_chan = chan;
      const int r = g_apps[i].exec(chan, data);
      t_chan = NULL;
      return r;
    }
  return -2;
}
int pbx_builtin_setvar_helper(struct ast_channel *chan, const char *name, const char *value) {
  if (!chan || chan->nvars >= 32) return -1;
  snprintf(chan->vars[chan->nvars][0], 512, "%s", name);
  snprintf(chan->vars[chan->nvars][1], 512, "%s", value ? value : "");
  chan->nvars++;
  return 0;
}
enum ast_channel_state ast_channel_state(const struct ast_channel *chan) { return chan->state; }
int ast_answer(struct ast_channel *chan) {
  chan->state = AST_STATE_UP, chan->answered++;
  return 0;
}
struct timeval ast_tvnow(void) {
  struct timeval tv = {0, 0};
  if (t_chan) tv.tv_sec = t_chan->clock_ms / 1000, tv.tv_usec = (t_chan->clock_ms % 1000) * 1000;
  else gettimeofday(&tv, NULL);
  return tv;
}
int ast_remaining_ms(struct timeval start, int max_ms) {
  if (max_ms < 0) return max_ms;
  const struct timeval now = ast_tvnow();
  const long el = (now.tv_sec - start.tv_sec) * 1000 + (now.tv_usec - start.tv_usec) / 1000;
  const long left = (long)max_ms - el;
  return left < 0 ? 0 : (int)left;
}
int ast_waitfor(struct ast_channel *chan, int ms) {
  /* audio keeps arriving until the script ends; then the wait times out (0) after consuming the rest of `ms` */
  if (chan->pos < chan->n_samples || (chan->hangup_at >= 0 && chan->pos >= chan->hangup_at)) return ms > 0 ? ms : 1;
  chan->clock_ms += ms > 0 ? ms : 0;
  return 0;
}
struct ast_frame *ast_read(struct ast_channel *chan) {
  if (chan->hangup_at >= 0 && chan->pos >= chan->hangup_at) return NULL;
  struct ast_frame *f = calloc(1, sizeof *f);
  chan->frames_read++;
  if (chan->control_every && chan->frames_read % chan->control_every == 0) {
    f->frametype = AST_FRAME_CONTROL;
    return f;
  }
  long n = chan->n_samples - chan->pos;
  if (n > chan->samples_per_frame) n = chan->samples_per_frame;
  if (chan->hangup_at >= 0 && chan->pos + n > chan->hangup_at) n = chan->hangup_at - chan->pos;
  f->frametype = AST_FRAME_VOICE;
  f->samples = (int)n, f->datalen = (int)n * 2;
  f->data.ptr = malloc((size_t)(n > 0 ? n : 1) * 2);
  memcpy(f->data.ptr, chan->pcm + chan->pos, (size_t)n * 2);
  chan->pos += n;
  chan->clock_ms += 1000L * chan->samples_per_frame / chan->rate;
  return f;
}
void ast_frfree(struct ast_frame *fr) {
  if (fr) free(fr->data.ptr), free(fr);
}

/* ------------------------------------------------------------------------------------------ "wav" file streams */
struct ast_filestream {
  char path[1024];
  int16_t *pcm;
  size_t n, cap;
  int rate;
};
struct ast_filestream *ast_writefile(const char *filename, const char *type, const char *comment, int flags, int check, mode_t mode) {
  (void)comment, (void)flags, (void)check, (void)mode;
  if (!filename || !type || strcmp(type, "wav") != 0) return NULL;
  struct ast_filestream *s = calloc(1, sizeof *s);
  snprintf(s->path, sizeof s->path, "%s.%s", filename, type);
  s->rate = t_chan ? t_chan->rate : 8000; /* the format of the writing channel (format_wav: 8 kHz slin; wav16: 16 kHz) */
  return s;
}
int ast_writestream(struct ast_filestream *fs, struct ast_frame *f) {
  if (!fs || !f || f->frametype != AST_FRAME_VOICE) return -1;
  if (fs->n + (size_t)f->samples > fs->cap) {
    fs->cap = (fs->n + (size_t)f->samples) * 2 + 4096;
    fs->pcm = realloc(fs->pcm, fs->cap * 2);
  }
  memcpy(fs->pcm + fs->n, f->data.ptr, (size_t)f->samples * 2);
  fs->n += (size_t)f->samples;
  return 0;
}
static void le32(unsigned char *p, uint32_t v) { p[0] = v & 255, p[1] = (v >> 8) & 255, p[2] = (v >> 16) & 255, p[3] = v >> 24; }
int ast_closestream(struct ast_filestream *s) {
  if (!s) return -1;
  FILE *f = fopen(s->path, "wb");
  int rc = -1;
  if (f) {
    unsigned char h[44];
    const uint32_t bytes = (uint32_t)(s->n * 2);
    memcpy(h, "RIFF", 4), le32(h + 4, 36 + bytes), memcpy(h + 8, "WAVEfmt ", 8), le32(h + 16, 16);
    h[20] = 1, h[21] = 0, h[22] = 1, h[23] = 0; /* PCM, mono */
    le32(h + 24, (uint32_t)s->rate), le32(h + 28, (uint32_t)s->rate * 2);
    h[32] = 2, h[33] = 0, h[34] = 16, h[35] = 0;
    memcpy(h + 36, "data", 4), le32(h + 40, bytes);
    rc = (fwrite(h, 1, 44, f) == 44 && fwrite(s->pcm, 2, s->n, f) == s->n) ? 0 : -1;
    fclose(f);
  }
  free(s->pcm), free(s);
  return rc;
}
int ast_filedelete(const char *filename, const char *fmt) {
  char p[1100];
  snprintf(p, sizeof p, "%s.%s", filename, fmt ? fmt : "wav");
  return unlink(p);
}

/* ------------------------------------------------------------------------------------------ module loader */
extern struct ast_module_info fake_ast_module_info; /* defined by src/app_tiresias.c through AST_MODULE_INFO */
int fake_module_load(void) { return fake_ast_module_info.load ? fake_ast_module_info.load() : -99; }
int fake_module_unload(void) { return fake_ast_module_info.unload ? fake_ast_module_info.unload() : -99; }
int fake_module_reload(void) { return fake_ast_module_info.reload ? fake_ast_module_info.reload() : -99; }
const char *fake_module_description(void) { return fake_ast_module_info.description; }
