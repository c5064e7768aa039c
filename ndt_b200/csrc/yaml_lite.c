/* yaml_lite.c -- libyaml's event API for ndt's scene files, without libyaml.
 *
 * See include/yaml_lite/yaml.h for scope.  Consumers: the reference's scene.c
 * (scene_read_yaml scene.c:2090, scene_write_yaml scene.c:1000,
 * scene_write_yaml_buffer scene.c:1045, scene_yaml_count_frames scene.c:2134)
 * and scenes/yaml.c, compiled unmodified against this API.
 *
 * PARSER: the whole input is parsed eagerly into an event array on the first
 * yaml_parser_parse call (scene files are a few MB at most); events are handed
 * out one per call.  A syntax error surfaces at the call that would have
 * returned the first event after it, like libyaml's.
 * EMITTER: the state machine of libyaml 0.2.5's emitter restricted to what an
 * anchor-less, tag-less event stream can reach, byte-identical output
 * (tests/test_yaml_lite.py compares with libyaml's own emitter).
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "yaml_lite/yaml.h"

const char *yaml_get_version_string(void) { return "yaml_lite-1 (libyaml 0.2.5 event API subset)"; }
void ylite_free(void *p) { free(p); }

/* ------------------------------------------------------------------ events */

static yaml_char_t *dup_n(const yaml_char_t *s, size_t n)
{
    yaml_char_t *d = (yaml_char_t *)malloc(n + 1);
    if (!d) return NULL;
    if (n) memcpy(d, s, n);
    d[n] = 0;
    return d;
}
static yaml_char_t *dup_z(const yaml_char_t *s) { return s ? dup_n(s, strlen((const char *)s)) : NULL; }

void yaml_event_delete(yaml_event_t *e)
{
    if (!e) return;
    switch (e->type) {
    case YAML_SCALAR_EVENT:
        free(e->data.scalar.anchor); free(e->data.scalar.tag); free(e->data.scalar.value); break;
    case YAML_SEQUENCE_START_EVENT:
        free(e->data.sequence_start.anchor); free(e->data.sequence_start.tag); break;
    case YAML_MAPPING_START_EVENT:
        free(e->data.mapping_start.anchor); free(e->data.mapping_start.tag); break;
    case YAML_ALIAS_EVENT:
        free(e->data.alias.anchor); break;
    default: break;
    }
    memset(e, 0, sizeof(*e));
}

int yaml_stream_start_event_initialize(yaml_event_t *e, yaml_encoding_t encoding)
{
    memset(e, 0, sizeof(*e));
    e->type = YAML_STREAM_START_EVENT;
    e->data.stream_start.encoding = encoding;
    return 1;
}
int yaml_stream_end_event_initialize(yaml_event_t *e)
{
    memset(e, 0, sizeof(*e));
    e->type = YAML_STREAM_END_EVENT;
    return 1;
}
int yaml_document_start_event_initialize(yaml_event_t *e, yaml_version_directive_t *vd,
        yaml_tag_directive_t *ts, yaml_tag_directive_t *te, int implicit)
{
    memset(e, 0, sizeof(*e));
    if (vd || (ts && ts != te)) return 0; /* directives: out of scope */
    e->type = YAML_DOCUMENT_START_EVENT;
    e->data.document_start.implicit = implicit;
    return 1;
}
int yaml_document_end_event_initialize(yaml_event_t *e, int implicit)
{
    memset(e, 0, sizeof(*e));
    e->type = YAML_DOCUMENT_END_EVENT;
    e->data.document_end.implicit = implicit;
    return 1;
}
int yaml_scalar_event_initialize(yaml_event_t *e, const yaml_char_t *anchor, const yaml_char_t *tag,
        const yaml_char_t *value, int length, int plain_implicit, int quoted_implicit, yaml_scalar_style_t style)
{
    memset(e, 0, sizeof(*e));
    if (!value) return 0;
    if (length < 0) length = (int)strlen((const char *)value);
    e->data.scalar.value = dup_n(value, (size_t)length);
    if (!e->data.scalar.value) return 0;
    e->type = YAML_SCALAR_EVENT;
    e->data.scalar.anchor = dup_z(anchor);
    e->data.scalar.tag = dup_z(tag);
    e->data.scalar.length = (size_t)length;
    e->data.scalar.plain_implicit = plain_implicit;
    e->data.scalar.quoted_implicit = quoted_implicit;
    e->data.scalar.style = style;
    return 1;
}
int yaml_sequence_start_event_initialize(yaml_event_t *e, const yaml_char_t *anchor, const yaml_char_t *tag,
        int implicit, yaml_sequence_style_t style)
{
    memset(e, 0, sizeof(*e));
    e->type = YAML_SEQUENCE_START_EVENT;
    e->data.sequence_start.anchor = dup_z(anchor);
    e->data.sequence_start.tag = dup_z(tag);
    e->data.sequence_start.implicit = implicit;
    e->data.sequence_start.style = style;
    return 1;
}
int yaml_sequence_end_event_initialize(yaml_event_t *e)
{
    memset(e, 0, sizeof(*e));
    e->type = YAML_SEQUENCE_END_EVENT;
    return 1;
}
int yaml_mapping_start_event_initialize(yaml_event_t *e, const yaml_char_t *anchor, const yaml_char_t *tag,
        int implicit, yaml_mapping_style_t style)
{
    memset(e, 0, sizeof(*e));
    e->type = YAML_MAPPING_START_EVENT;
    e->data.mapping_start.anchor = dup_z(anchor);
    e->data.mapping_start.tag = dup_z(tag);
    e->data.mapping_start.implicit = implicit;
    e->data.mapping_start.style = style;
    return 1;
}
int yaml_mapping_end_event_initialize(yaml_event_t *e)
{
    memset(e, 0, sizeof(*e));
    e->type = YAML_MAPPING_END_EVENT;
    return 1;
}

/* ------------------------------------------------------------------ parser */

struct ylite_parser_impl {
    unsigned char *own;          /* file contents (owned) */
    const unsigned char *s;      /* input */
    size_t n, pos, line, col;
    yaml_event_t *ev;            /* parsed events */
    size_t nev, cap, next;
    int parsed;
    int failed;                  /* events up to nev are valid, then the error */
    yaml_error_type_t err;
    const char *problem;
    yaml_mark_t problem_mark;
    int done;                    /* STREAM-END handed out */
    unsigned char *sc;           /* scalar scratch */
    size_t sc_n, sc_cap;
    size_t sc_floor;             /* scratch bytes produced by escapes: not trimmed when a line folds */
    size_t plain_indent;         /* block context: a plain scalar continues on lines indented at least this far */
};
typedef struct ylite_parser_impl P;

int yaml_parser_initialize(yaml_parser_t *parser)
{
    memset(parser, 0, sizeof(*parser));
    parser->impl = (P *)calloc(1, sizeof(P));
    if (!parser->impl) { parser->error = YAML_MEMORY_ERROR; return 0; }
    return 1;
}

void yaml_parser_delete(yaml_parser_t *parser)
{
    P *p = parser->impl;
    if (p) {
        for (size_t i = p->next; i < p->nev; ++i) yaml_event_delete(&p->ev[i]);
        free(p->ev); free(p->own); free(p->sc); free(p);
    }
    memset(parser, 0, sizeof(*parser));
}

void yaml_parser_set_input_string(yaml_parser_t *parser, const unsigned char *input, size_t size)
{
    P *p = parser->impl;
    p->s = input; p->n = size;
}

void yaml_parser_set_input_file(yaml_parser_t *parser, FILE *file)
{
    P *p = parser->impl;
    size_t cap = 1 << 16, n = 0;
    unsigned char *buf = (unsigned char *)malloc(cap);
    while (buf) {
        size_t r = fread(buf + n, 1, cap - n, file);
        n += r;
        if (r == 0) break;
        if (n == cap) {
            unsigned char *t = (unsigned char *)realloc(buf, cap *= 2);
            if (!t) { free(buf); buf = NULL; }
            else buf = t;
        }
    }
    p->own = buf; p->s = buf; p->n = buf ? n : 0;
}

static int p_fail(P *p, yaml_error_type_t code, const char *problem)
{
    if (!p->failed) {
        p->failed = 1; p->err = code; p->problem = problem;
        p->problem_mark.index = p->pos; p->problem_mark.line = p->line; p->problem_mark.column = p->col;
    }
    return 0;
}

static yaml_event_t *p_push(P *p, yaml_event_type_t type)
{
    if (p->nev == p->cap) {
        size_t nc = p->cap ? p->cap * 2 : 256;
        yaml_event_t *t = (yaml_event_t *)realloc(p->ev, nc * sizeof(yaml_event_t));
        if (!t) { p_fail(p, YAML_MEMORY_ERROR, "out of memory"); return NULL; }
        p->ev = t; p->cap = nc;
    }
    yaml_event_t *e = &p->ev[p->nev++];
    memset(e, 0, sizeof(*e));
    e->type = type;
    e->start_mark.index = e->end_mark.index = p->pos;
    e->start_mark.line = e->end_mark.line = p->line;
    e->start_mark.column = e->end_mark.column = p->col;
    return e;
}

static int p_scalar(P *p, const unsigned char *v, size_t n, yaml_scalar_style_t style)
{
    yaml_event_t *e = p_push(p, YAML_SCALAR_EVENT);
    if (!e) return 0;
    e->data.scalar.value = dup_n(v, n);
    if (!e->data.scalar.value) { p->nev--; return p_fail(p, YAML_MEMORY_ERROR, "out of memory"); }
    e->data.scalar.length = n;
    e->data.scalar.style = style;
    e->data.scalar.plain_implicit = (style == YAML_PLAIN_SCALAR_STYLE);
    e->data.scalar.quoted_implicit = (style != YAML_PLAIN_SCALAR_STYLE);
    return 1;
}
static int p_empty(P *p) { return p_scalar(p, (const unsigned char *)"", 0, YAML_PLAIN_SCALAR_STYLE); }

static int pk(const P *p, size_t k) { return p->pos + k < p->n ? p->s[p->pos + k] : 0; }
static int at_eof(const P *p) { return p->pos >= p->n || p->s[p->pos] == 0; }
static int is_brk(int c) { return c == '\n' || c == '\r'; }
static int is_blank(int c) { return c == ' ' || c == '\t'; }
static int is_blankz(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == 0; }
static int is_flowind(int c) { return c == ',' || c == '[' || c == ']' || c == '{' || c == '}'; }

static void p_adv(P *p) { p->pos++; p->col++; }
static void p_brk(P *p)
{
    if (pk(p, 0) == '\r' && pk(p, 1) == '\n') p->pos += 2; else p->pos++;
    p->line++; p->col = 0;
}

/* blanks on this line, then a comment if one starts here */
static void skip_line_tail(P *p)
{
    int prev_blank = (p->col == 0) || (p->pos > 0 && is_blankz(p->s[p->pos - 1]));
    while (is_blank(pk(p, 0))) { p_adv(p); prev_blank = 1; }
    if (pk(p, 0) == '#' && prev_blank) while (!at_eof(p) && !is_brk(pk(p, 0))) p_adv(p);
}
/* to the next content character, over blanks, comments and line breaks */
static int skip_to_content(P *p)
{
    for (;;) {
        size_t line_start = (p->col == 0);
        size_t c0 = p->pos;
        skip_line_tail(p);
        if (line_start && is_brk(pk(p, 0)) == 0 && !at_eof(p)) {
            /* indentation must be spaces */
            for (size_t i = c0; i < p->pos; ++i)
                if (p->s[i] == '\t' && pk(p, 0) != '#')
                    return p_fail(p, YAML_SCANNER_ERROR, "found character that cannot start any token (tab in indentation)");
        }
        if (is_brk(pk(p, 0))) { p_brk(p); continue; }
        return 1;
    }
}
static int at_marker(const P *p)
{
    if (p->col != 0) return 0;
    int a = pk(p, 0);
    return (a == '-' || a == '.') && pk(p, 1) == a && pk(p, 2) == a && is_blankz(pk(p, 3));
}
static int at_line_end(const P *p) { return at_eof(p) || is_brk(pk(p, 0)); }

static int sc_put(P *p, int c)
{
    if (p->sc_n + 1 >= p->sc_cap) {
        size_t nc = p->sc_cap ? p->sc_cap * 2 : 256;
        unsigned char *t = (unsigned char *)realloc(p->sc, nc);
        if (!t) return p_fail(p, YAML_MEMORY_ERROR, "out of memory");
        p->sc = t; p->sc_cap = nc;
    }
    p->sc[p->sc_n++] = (unsigned char)c;
    return 1;
}
static int sc_put_utf8(P *p, unsigned v)
{
    if (v <= 0x7F) return sc_put(p, (int)v);
    if (v <= 0x7FF) return sc_put(p, 0xC0 | (v >> 6)) && sc_put(p, 0x80 | (v & 0x3F));
    if (v <= 0xFFFF) return sc_put(p, 0xE0 | (v >> 12)) && sc_put(p, 0x80 | ((v >> 6) & 0x3F)) && sc_put(p, 0x80 | (v & 0x3F));
    return sc_put(p, 0xF0 | (v >> 18)) && sc_put(p, 0x80 | ((v >> 12) & 0x3F))
        && sc_put(p, 0x80 | ((v >> 6) & 0x3F)) && sc_put(p, 0x80 | (v & 0x3F));
}

/* line folding inside a quoted scalar: positioned at a break */
static int fold_breaks(P *p)
{
    int breaks = 0;
    /* trailing blanks before the break are dropped */
    while (p->sc_n > p->sc_floor && is_blank(p->sc[p->sc_n - 1])) p->sc_n--;
    for (;;) {
        while (is_blank(pk(p, 0))) p_adv(p);
        if (is_brk(pk(p, 0))) { p_brk(p); breaks++; continue; }
        break;
    }
    if (at_marker(p)) return p_fail(p, YAML_SCANNER_ERROR, "found unexpected document indicator");
    if (at_eof(p)) return p_fail(p, YAML_SCANNER_ERROR, "found unexpected end of stream");
    if (breaks == 1) return sc_put(p, ' ');
    for (int i = 1; i < breaks; ++i) if (!sc_put(p, '\n')) return 0;
    return 1;
}

static int hexval(int c)
{
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

/* quoted scalar into the scratch buffer; positioned at the opening quote */
static int scan_quoted(P *p, int *style)
{
    int q = pk(p, 0);
    *style = (q == '\'') ? YAML_SINGLE_QUOTED_SCALAR_STYLE : YAML_DOUBLE_QUOTED_SCALAR_STYLE;
    p->sc_n = 0; p->sc_floor = 0;
    p_adv(p);
    for (;;) {
        int c = pk(p, 0);
        if (at_eof(p)) return p_fail(p, YAML_SCANNER_ERROR, "found unexpected end of stream");
        if (is_brk(c)) { if (!fold_breaks(p)) return 0; continue; }
        if (q == '\'') {
            if (c == '\'') {
                if (pk(p, 1) == '\'') { if (!sc_put(p, '\'')) return 0; p_adv(p); p_adv(p); continue; }
                p_adv(p); return 1;
            }
        } else {
            if (c == '"') { p_adv(p); return 1; }
            if (c == '\\') {
                int e = pk(p, 1), len = 0;
                if (is_brk(e)) {           /* escaped line break: join, drop leading blanks */
                    p_adv(p); p_brk(p);
                    for (;;) {
                        while (is_blank(pk(p, 0))) p_adv(p);
                        if (is_brk(pk(p, 0))) { p_brk(p); if (!sc_put(p, '\n')) return 0; continue; }
                        break;
                    }
                    continue;
                }
                p_adv(p); p_adv(p);
                switch (e) {
                case '0': sc_put(p, 0); break;
                case 'a': sc_put(p, 7); break;
                case 'b': sc_put(p, 8); break;
                case 't': case '\t': sc_put(p, 9); break;
                case 'n': sc_put(p, 10); break;
                case 'v': sc_put(p, 11); break;
                case 'f': sc_put(p, 12); break;
                case 'r': sc_put(p, 13); break;
                case 'e': sc_put(p, 27); break;
                case ' ': sc_put(p, ' '); break;
                case '"': sc_put(p, '"'); break;
                case '/': sc_put(p, '/'); break;
                case '\\': sc_put(p, '\\'); break;
                case 'N': sc_put_utf8(p, 0x85); break;
                case '_': sc_put_utf8(p, 0xA0); break;
                case 'L': sc_put_utf8(p, 0x2028); break;
                case 'P': sc_put_utf8(p, 0x2029); break;
                case 'x': len = 2; break;
                case 'u': len = 4; break;
                case 'U': len = 8; break;
                default: return p_fail(p, YAML_SCANNER_ERROR, "found unknown escape character");
                }
                if (len) {
                    unsigned v = 0;
                    for (int i = 0; i < len; ++i) {
                        int h = hexval(pk(p, 0));
                        if (h < 0) return p_fail(p, YAML_SCANNER_ERROR, "did not find expected hexdecimal number");
                        v = (v << 4) | (unsigned)h; p_adv(p);
                    }
                    if ((v >= 0xD800 && v <= 0xDFFF) || v > 0x10FFFF)
                        return p_fail(p, YAML_SCANNER_ERROR, "found invalid Unicode character escape code");
                    if (!sc_put_utf8(p, v)) return 0;
                }
                if (p->failed) return 0;
                p->sc_floor = p->sc_n;
                continue;
            }
        }
        if (!sc_put(p, c)) return 0;
        p_adv(p);
    }
}

/* plain scalar (one line) into the scratch buffer */
static int scan_plain(P *p, int flow)
{
    p->sc_n = 0;
    int c = pk(p, 0);
    if (c == '&' || c == '*' || c == '!' || c == '|' || c == '>' || c == '%' || c == '@' || c == '`')
        return p_fail(p, YAML_SCANNER_ERROR, "anchors, aliases, tags, block scalars and reserved indicators are not supported");
    if ((c == '?' || c == '-') && is_blankz(pk(p, 1)))
        return p_fail(p, YAML_SCANNER_ERROR, c == '?' ? "a complex mapping key is not allowed in this context"
                                                      : "block sequence entries are not allowed in this context");
    if (c == ':' && is_blankz(pk(p, 1)))
        return p_fail(p, YAML_SCANNER_ERROR, "mapping values are not allowed in this context");
    if (is_flowind(c) || c == '#')
        return p_fail(p, YAML_SCANNER_ERROR, "found character that cannot start any token");
    for (;;) {
        for (;;) {
            c = pk(p, 0);
            if (at_eof(p) || is_brk(c)) break;
            if (c == ':' && is_blankz(pk(p, 1))) break;
            if (flow && c == ':' && is_flowind(pk(p, 1)))
                return p_fail(p, YAML_SCANNER_ERROR, "found unexpected ':'");
            if (flow && is_flowind(c)) break;
            if (c == '#' && p->sc_n && is_blank(p->sc[p->sc_n - 1])) break;
            if (!sc_put(p, c)) return 0;
            p_adv(p);
        }
        while (p->sc_n && is_blank(p->sc[p->sc_n - 1])) p->sc_n--;
        if (!is_brk(pk(p, 0)) || at_eof(p)) break;
        /* a plain scalar continues on the next non-empty line when that line is content indented far enough
         * (block context) and does not start with something that ends a plain scalar: one break folds into a
         * space, k breaks into k-1 line feeds (libyaml: yaml_parser_scan_plain_scalar) */
        const size_t s_pos = p->pos, s_line = p->line, s_col = p->col;
        int breaks = 0, tab_indent = 0;
        for (;;) {
            if (is_brk(pk(p, 0))) { p_brk(p); breaks++; continue; }
            if (pk(p, 0) == ' ') { p_adv(p); continue; }
            if (pk(p, 0) == '\t') { tab_indent = 1; p_adv(p); continue; }
            break;
        }
        c = pk(p, 0);
        int cont = !at_eof(p) && !at_marker(p) && c != '#';
        if (cont && !flow && p->col < p->plain_indent) cont = 0;
        if (cont && flow && (c == ',' || c == ']' || c == '}' ||
                             (c == ':' && (is_blankz(pk(p, 1)) || is_flowind(pk(p, 1)))))) cont = 0;
        if (cont && !flow && tab_indent && p->col < p->plain_indent) cont = 0;
        if (!cont) { p->pos = s_pos; p->line = s_line; p->col = s_col; break; }
        if (c == ':' && is_blankz(pk(p, 1)))
            return p_fail(p, YAML_SCANNER_ERROR, "mapping values are not allowed in this context");
        if (breaks == 1) { if (!sc_put(p, ' ')) return 0; }
        else for (int i = 1; i < breaks; ++i) if (!sc_put(p, '\n')) return 0;
    }
    return 1;
}

static int scan_scalar(P *p, int flow, int *style)
{
    int c = pk(p, 0);
    if (c == '\'' || c == '"') return scan_quoted(p, style);
    *style = YAML_PLAIN_SCALAR_STYLE;
    return scan_plain(p, flow);
}

static int parse_flow_node(P *p);

static int skip_flow_space(P *p)
{
    if (!skip_to_content(p)) return 0;
    if (at_marker(p)) return p_fail(p, YAML_SCANNER_ERROR, "found unexpected document indicator");
    if (at_eof(p)) return p_fail(p, YAML_PARSER_ERROR, "did not find expected ',' or closing bracket");
    return 1;
}

/* ':' that separates a flow key from its value */
static int flow_colon(const P *p, int after_quoted)
{
    if (pk(p, 0) != ':') return 0;
    int d = pk(p, 1);
    return is_blankz(d) || is_flowind(d) || after_quoted;
}

static int parse_flow_seq(P *p)
{
    yaml_event_t *e = p_push(p, YAML_SEQUENCE_START_EVENT);
    if (!e) return 0;
    e->data.sequence_start.implicit = 1;
    e->data.sequence_start.style = YAML_FLOW_SEQUENCE_STYLE;
    p_adv(p);
    for (;;) {
        if (!skip_flow_space(p)) return 0;
        if (pk(p, 0) == ']') { p_adv(p); break; }
        size_t at = p->nev;
        const size_t key_line = p->line;
        int q = (pk(p, 0) == '\'' || pk(p, 0) == '"');
        if (!parse_flow_node(p)) return 0;
        const int one_line = p->line == key_line;
        if (!skip_flow_space(p)) return 0;
        if (flow_colon(p, q) && p->nev == at + 1 && !one_line)
            return p_fail(p, YAML_SCANNER_ERROR, "mapping values are not allowed in this context (a simple key may not span lines)");
        if (flow_colon(p, q) && p->nev == at + 1) {
            /* single-pair mapping inside a flow sequence: [a: b] */
            if (!p_push(p, YAML_NO_EVENT)) return 0;
            memmove(&p->ev[at + 1], &p->ev[at], (p->nev - 1 - at) * sizeof(yaml_event_t));
            memset(&p->ev[at], 0, sizeof(yaml_event_t));
            p->ev[at].type = YAML_MAPPING_START_EVENT;
            p->ev[at].data.mapping_start.implicit = 1;
            p->ev[at].data.mapping_start.style = YAML_FLOW_MAPPING_STYLE;
            p->ev[at].start_mark = p->ev[at].end_mark = p->ev[at + 1].start_mark;
            p_adv(p);
            if (!skip_flow_space(p)) return 0;
            if (pk(p, 0) == ',' || pk(p, 0) == ']') { if (!p_empty(p)) return 0; }
            else if (!parse_flow_node(p)) return 0;
            if (!p_push(p, YAML_MAPPING_END_EVENT)) return 0;
            if (!skip_flow_space(p)) return 0;
        }
        if (pk(p, 0) == ',') { p_adv(p); continue; }
        if (pk(p, 0) == ']') { p_adv(p); break; }
        return p_fail(p, YAML_PARSER_ERROR, "did not find expected ',' or ']'");
    }
    return p_push(p, YAML_SEQUENCE_END_EVENT) != NULL;
}

static int parse_flow_map(P *p)
{
    yaml_event_t *e = p_push(p, YAML_MAPPING_START_EVENT);
    if (!e) return 0;
    e->data.mapping_start.implicit = 1;
    e->data.mapping_start.style = YAML_FLOW_MAPPING_STYLE;
    p_adv(p);
    for (;;) {
        if (!skip_flow_space(p)) return 0;
        if (pk(p, 0) == '}') { p_adv(p); break; }
        int complex_key = 0;
        if (pk(p, 0) == '?' && is_blankz(pk(p, 1))) {        /* {? key : value} */
            complex_key = 1;
            p_adv(p);
            if (!skip_flow_space(p)) return 0;
        }
        int q = (pk(p, 0) == '\'' || pk(p, 0) == '"') || complex_key;
        if ((pk(p, 0) == ':' && (is_blankz(pk(p, 1)) || is_flowind(pk(p, 1)))) ||
            (complex_key && (pk(p, 0) == ',' || pk(p, 0) == '}'))) {
            if (!p_empty(p)) return 0;       /* {: v}, {? } */
        } else if (complex_key) {
            if (!parse_flow_node(p)) return 0;
        } else {
            const size_t key_line = p->line;
            if (!parse_flow_node(p)) return 0;
            if (p->line != key_line && p->ev[p->nev - 1].type == YAML_SCALAR_EVENT) {
                /* a key that spans lines is only legal when no ':' follows it */
                const size_t s_pos = p->pos, s_line = p->line, s_col = p->col;
                if (!skip_flow_space(p)) return 0;
                if (flow_colon(p, q))
                    return p_fail(p, YAML_SCANNER_ERROR, "mapping values are not allowed in this context (a simple key may not span lines)");
                p->pos = s_pos; p->line = s_line; p->col = s_col;
            }
        }
        if (!skip_flow_space(p)) return 0;
        if (flow_colon(p, q)) {
            p_adv(p);
            if (!skip_flow_space(p)) return 0;
            if (pk(p, 0) == ',' || pk(p, 0) == '}') { if (!p_empty(p)) return 0; }
            else if (!parse_flow_node(p)) return 0;
            if (!skip_flow_space(p)) return 0;
        } else if (!p_empty(p)) return 0;
        if (pk(p, 0) == ',') { p_adv(p); continue; }
        if (pk(p, 0) == '}') { p_adv(p); break; }
        return p_fail(p, YAML_PARSER_ERROR, "did not find expected ',' or '}'");
    }
    return p_push(p, YAML_MAPPING_END_EVENT) != NULL;
}

static int parse_flow_node(P *p)
{
    int c = pk(p, 0), style;
    if (c == '[') return parse_flow_seq(p);
    if (c == '{') return parse_flow_map(p);
    if (!scan_scalar(p, 1, &style)) return 0;
    return p_scalar(p, p->sc, p->sc_n, (yaml_scalar_style_t)style);
}

static int parse_block_node(P *p);
static int parse_block_seq(P *p, size_t n, int indentless);

/* after a node that ends on this line: only a comment may follow */
static int expect_line_end(P *p)
{
    skip_line_tail(p);
    if (!at_line_end(p)) {
        if (pk(p, 0) == ':' && is_blankz(pk(p, 1)))
            return p_fail(p, YAML_SCANNER_ERROR, "mapping values are not allowed in this context");
        return p_fail(p, YAML_PARSER_ERROR, "did not find expected end of line");
    }
    return 1;
}

/* value of a block mapping key at indent n; positioned just after ':' */
static int parse_map_value(P *p, size_t n)
{
    p->plain_indent = n + 1;
    skip_line_tail(p);
    if (at_line_end(p)) {
        if (!skip_to_content(p)) return 0;
        if (at_eof(p) || at_marker(p)) return p_empty(p);
        if (p->col > n) return parse_block_node(p);
        if (p->col == n && pk(p, 0) == '-' && is_blankz(pk(p, 1))) return parse_block_seq(p, n, 1);
        return p_empty(p);
    }
    int c = pk(p, 0), style;
    if (c == '[') return parse_flow_seq(p) && expect_line_end(p);
    if (c == '{') return parse_flow_map(p) && expect_line_end(p);
    if (!scan_scalar(p, 0, &style)) return 0;
    if (!p_scalar(p, p->sc, p->sc_n, (yaml_scalar_style_t)style)) return 0;
    return expect_line_end(p);
}

/* block mapping at indent n whose first key is in the scratch buffer; positioned at its ':' */
static int parse_block_map(P *p, size_t n, int style)
{
    yaml_event_t *e = p_push(p, YAML_MAPPING_START_EVENT);
    if (!e) return 0;
    e->data.mapping_start.implicit = 1;
    e->data.mapping_start.style = YAML_BLOCK_MAPPING_STYLE;
    e->start_mark.column = e->end_mark.column = n;
    for (;;) {
        if (style >= 0) {
            if (!p_scalar(p, p->sc, p->sc_n, (yaml_scalar_style_t)style)) return 0;
            p_adv(p);                               /* ':' */
            if (!parse_map_value(p, n)) return 0;
        } else {
            /* complex key: "? key" and, on a line of its own at the same indent, ": value" */
            p->plain_indent = n + 1;
            p_adv(p);                               /* '?' */
            skip_line_tail(p);
            if (at_line_end(p)) {
                if (!skip_to_content(p)) return 0;
                if (!at_eof(p) && !at_marker(p) && p->col > n) { if (!parse_block_node(p)) return 0; }
                else if (!p_empty(p)) return 0;
            } else if (!parse_block_node(p)) return 0;
            if (!skip_to_content(p)) return 0;
            if (!at_eof(p) && !at_marker(p) && p->col == n && pk(p, 0) == ':' && is_blankz(pk(p, 1))) {
                p->plain_indent = n + 1;
                p_adv(p);                           /* ':' */
                skip_line_tail(p);
                if (at_line_end(p)) {
                    if (!skip_to_content(p)) return 0;
                    if (!at_eof(p) && !at_marker(p) && p->col > n) { if (!parse_block_node(p)) return 0; }
                    else if (!at_eof(p) && !at_marker(p) && p->col == n && pk(p, 0) == '-' && is_blankz(pk(p, 1))) {
                        if (!parse_block_seq(p, n, 1)) return 0;
                    } else if (!p_empty(p)) return 0;
                } else if (!parse_block_node(p)) return 0;
            } else if (!p_empty(p)) return 0;
        }
        if (!skip_to_content(p)) return 0;
        if (at_eof(p) || at_marker(p) || p->col < n) break;
        if (p->col > n) return p_fail(p, YAML_PARSER_ERROR, "did not find expected key (bad indentation of a mapping entry)");
        if (pk(p, 0) == '?' && is_blankz(pk(p, 1))) { style = -1; continue; }
        if (pk(p, 0) == '[' || pk(p, 0) == '{')
            return p_fail(p, YAML_PARSER_ERROR, "flow collections as mapping keys are not supported");
        const size_t key_line = p->line;
        p->plain_indent = n + 1;
        if (!scan_scalar(p, 0, &style)) return 0;
        while (is_blank(pk(p, 0))) p_adv(p);
        if (!(pk(p, 0) == ':' && is_blankz(pk(p, 1))) || p->line != key_line)
            return p_fail(p, YAML_SCANNER_ERROR, "could not find expected ':'");
    }
    return p_push(p, YAML_MAPPING_END_EVENT) != NULL;
}

/* block sequence at indent n; positioned at its first '-' */
static int parse_block_seq(P *p, size_t n, int indentless)
{
    yaml_event_t *e = p_push(p, YAML_SEQUENCE_START_EVENT);
    if (!e) return 0;
    e->data.sequence_start.implicit = 1;
    e->data.sequence_start.style = YAML_BLOCK_SEQUENCE_STYLE;
    for (;;) {
        p->plain_indent = n + 1;
        p_adv(p);                                   /* '-' */
        skip_line_tail(p);
        if (at_line_end(p)) {
            if (!skip_to_content(p)) return 0;
            if (!at_eof(p) && !at_marker(p) && p->col > n) { if (!parse_block_node(p)) return 0; }
            else if (!p_empty(p)) return 0;
        } else if (!parse_block_node(p)) return 0;
        if (!skip_to_content(p)) return 0;
        if (at_eof(p) || at_marker(p) || p->col < n) break;
        if (p->col > n) return p_fail(p, YAML_PARSER_ERROR, "did not find expected '-' indicator (bad indentation of a sequence entry)");
        if (pk(p, 0) == '-' && is_blankz(pk(p, 1))) continue;
        if (indentless) break;
        return p_fail(p, YAML_PARSER_ERROR, "did not find expected '-' indicator");
    }
    return p_push(p, YAML_SEQUENCE_END_EVENT) != NULL;
}

/* a node that starts at the current position in block context */
static int parse_block_node(P *p)
{
    int c = pk(p, 0), style;
    size_t col = p->col;
    if (c == '-' && is_blankz(pk(p, 1))) return parse_block_seq(p, col, 0);
    if (c == '?' && is_blankz(pk(p, 1))) return parse_block_map(p, col, -1);
    if (c == '[' || c == '{') {
        if (!(c == '[' ? parse_flow_seq(p) : parse_flow_map(p))) return 0;
        skip_line_tail(p);
        if (pk(p, 0) == ':' && is_blankz(pk(p, 1)))
            return p_fail(p, YAML_PARSER_ERROR, "flow collections as mapping keys are not supported");
        return expect_line_end(p);
    }
    size_t line = p->line;
    if (!scan_scalar(p, 0, &style)) return 0;
    while (is_blank(pk(p, 0))) p_adv(p);
    if (pk(p, 0) == ':' && is_blankz(pk(p, 1))) {
        if (p->line != line) return p_fail(p, YAML_SCANNER_ERROR, "a simple key may not span lines");
        return parse_block_map(p, col, style);
    }
    if (!p_scalar(p, p->sc, p->sc_n, (yaml_scalar_style_t)style)) return 0;
    return expect_line_end(p);
}

static void parse_stream(P *p)
{
    p->parsed = 1;
    if (p->n >= 3 && p->s[0] == 0xEF && p->s[1] == 0xBB && p->s[2] == 0xBF) p->pos = 3;   /* UTF-8 BOM */
    yaml_event_t *e = p_push(p, YAML_STREAM_START_EVENT);
    if (!e) return;
    e->data.stream_start.encoding = YAML_UTF8_ENCODING;
    int first = 1;
    for (;;) {
        if (!skip_to_content(p)) return;
        /* stray "..." between documents */
        while (at_marker(p) && pk(p, 0) == '.') {
            if (first) { p_fail(p, YAML_PARSER_ERROR, "did not find expected <document start>"); return; }
            p->pos += 3; p->col += 3;
            if (!expect_line_end(p) || !skip_to_content(p)) return;
        }
        if (at_eof(p)) break;
        if (pk(p, 0) == '%' && p->col == 0) { p_fail(p, YAML_SCANNER_ERROR, "directives are not supported"); return; }
        int explicit_start = at_marker(p);
        if (!explicit_start && !first) {      /* libyaml 0.2.5: only the first document may be implicit */
            p_fail(p, YAML_PARSER_ERROR, "did not find expected <document start>"); return;
        }
        e = p_push(p, YAML_DOCUMENT_START_EVENT);
        if (!e) return;
        e->data.document_start.implicit = !explicit_start;
        first = 0;
        int have_node = 0;
        p->plain_indent = 0;
        if (explicit_start) {
            p->pos += 3; p->col += 3;
            skip_line_tail(p);
            if (!at_line_end(p)) {
                /* content on the "---" line: a flow collection or a scalar */
                int c = pk(p, 0), style;
                if (c == '[') { if (!parse_flow_seq(p) || !expect_line_end(p)) return; }
                else if (c == '{') { if (!parse_flow_map(p) || !expect_line_end(p)) return; }
                else {
                    if (!scan_scalar(p, 0, &style) || !p_scalar(p, p->sc, p->sc_n, (yaml_scalar_style_t)style)
                        || !expect_line_end(p)) return;
                }
                have_node = 1;
            }
        }
        if (!have_node) {
            if (!skip_to_content(p)) return;
            if (at_eof(p) || at_marker(p)) { if (!p_empty(p)) return; }
            else if (!parse_block_node(p)) return;
        }
        if (!skip_to_content(p)) return;
        int explicit_end = 0;
        if (at_marker(p) && pk(p, 0) == '.') {
            explicit_end = 1;
            p->pos += 3; p->col += 3;
            if (!expect_line_end(p)) return;
        } else if (!at_eof(p) && !at_marker(p)) {
            p_fail(p, YAML_PARSER_ERROR, "did not find expected <document start>"); return;
        }
        e = p_push(p, YAML_DOCUMENT_END_EVENT);
        if (!e) return;
        e->data.document_end.implicit = !explicit_end;
    }
    p_push(p, YAML_STREAM_END_EVENT);
}

int yaml_parser_parse(yaml_parser_t *parser, yaml_event_t *event)
{
    P *p = parser->impl;
    memset(event, 0, sizeof(*event));
    if (!p) { parser->error = YAML_MEMORY_ERROR; return 0; }
    if (parser->error) return 0;
    if (p->done) return 1;
    if (!p->parsed) {
        if (!p->s) { p->s = (const unsigned char *)""; p->n = 0; }
        parse_stream(p);
    }
    if (p->next < p->nev) {
        *event = p->ev[p->next];
        memset(&p->ev[p->next], 0, sizeof(yaml_event_t));
        p->next++;
        if (event->type == YAML_STREAM_END_EVENT) p->done = 1;
        return 1;
    }
    if (p->failed) {
        parser->error = p->err;
        parser->problem = p->problem;
        parser->problem_mark = p->problem_mark;
        parser->problem_offset = p->problem_mark.index;
        return 0;
    }
    p->done = 1;
    return 1;
}

/* ----------------------------------------------------------------- emitter */

enum {
    ES_STREAM_START, ES_FIRST_DOCUMENT_START, ES_DOCUMENT_START, ES_DOCUMENT_CONTENT, ES_DOCUMENT_END,
    ES_FLOW_SEQ_FIRST_ITEM, ES_FLOW_SEQ_ITEM, ES_FLOW_MAP_FIRST_KEY, ES_FLOW_MAP_KEY,
    ES_FLOW_MAP_SIMPLE_VALUE, ES_FLOW_MAP_VALUE,
    ES_BLOCK_SEQ_FIRST_ITEM, ES_BLOCK_SEQ_ITEM, ES_BLOCK_MAP_FIRST_KEY, ES_BLOCK_MAP_KEY,
    ES_BLOCK_MAP_SIMPLE_VALUE, ES_BLOCK_MAP_VALUE, ES_END
};

struct ylite_emitter_impl {
    yaml_write_handler_t *handler;
    void *handler_data;
    unsigned char *out_buf; size_t out_size; size_t *out_written;   /* string output */
    FILE *file;
    unsigned char *buf; size_t len, cap;                             /* pending output */
    int best_indent, best_width;
    int state;
    int *states; size_t nstates, cstates;
    int *indents; size_t nindents, cindents;
    yaml_event_t *q; size_t qhead, qtail, qcap;
    int indent, flow_level;
    int root_context, sequence_context, mapping_context, simple_key_context;
    int line, column, whitespace, indention, open_ended;
    struct {
        const unsigned char *value; size_t length;
        int multiline, flow_plain_allowed, block_plain_allowed, single_quoted_allowed, block_allowed;
        yaml_scalar_style_t style;
        int bang;
    } sd;
};
typedef struct ylite_emitter_impl E;

static int e_fail(yaml_emitter_t *em, yaml_error_type_t code, const char *problem)
{
    em->error = code; em->problem = problem;
    return 0;
}

int yaml_emitter_initialize(yaml_emitter_t *em)
{
    memset(em, 0, sizeof(*em));
    em->impl = (E *)calloc(1, sizeof(E));
    if (!em->impl) { em->error = YAML_MEMORY_ERROR; return 0; }
    em->impl->best_indent = 2;
    em->impl->best_width = 80;
    em->impl->state = ES_STREAM_START;
    return 1;
}

void yaml_emitter_delete(yaml_emitter_t *em)
{
    E *e = em->impl;
    if (e) {
        for (size_t i = e->qhead; i < e->qtail; ++i) yaml_event_delete(&e->q[i]);
        free(e->q); free(e->states); free(e->indents); free(e->buf); free(e);
    }
    memset(em, 0, sizeof(*em));
}

static int string_write_handler(void *data, unsigned char *buffer, size_t size)
{
    E *e = (E *)data;
    size_t room = e->out_size - *e->out_written;
    if (room < size) {
        memcpy(e->out_buf + *e->out_written, buffer, room);
        *e->out_written = e->out_size;
        return 0;
    }
    memcpy(e->out_buf + *e->out_written, buffer, size);
    *e->out_written += size;
    return 1;
}
static int file_write_handler(void *data, unsigned char *buffer, size_t size)
{
    E *e = (E *)data;
    return fwrite(buffer, 1, size, e->file) == size;
}

void yaml_emitter_set_output_string(yaml_emitter_t *em, unsigned char *output, size_t size, size_t *size_written)
{
    E *e = em->impl;
    e->handler = string_write_handler; e->handler_data = e;
    e->out_buf = output; e->out_size = size; e->out_written = size_written;
    *size_written = 0;
}
void yaml_emitter_set_output_file(yaml_emitter_t *em, FILE *file)
{
    E *e = em->impl;
    e->handler = file_write_handler; e->handler_data = e; e->file = file;
}
void yaml_emitter_set_output(yaml_emitter_t *em, yaml_write_handler_t *handler, void *data)
{
    em->impl->handler = handler; em->impl->handler_data = data;
}
void yaml_emitter_set_width(yaml_emitter_t *em, int width) { em->impl->best_width = width; }
void yaml_emitter_set_indent(yaml_emitter_t *em, int indent) { em->impl->best_indent = indent; }

int yaml_emitter_flush(yaml_emitter_t *em)
{
    E *e = em->impl;
    if (!e || !e->handler) return e_fail(em, YAML_WRITER_ERROR, "no output set");
    if (!e->len) return 1;
    int ok = e->handler(e->handler_data, e->buf, e->len);
    e->len = 0;
    return ok ? 1 : e_fail(em, YAML_WRITER_ERROR, "write error");
}

static int w_put(yaml_emitter_t *em, int c)
{
    E *e = em->impl;
    if (e->len == e->cap) {
        size_t nc = e->cap ? e->cap * 2 : 1 << 14;
        unsigned char *t = (unsigned char *)realloc(e->buf, nc);
        if (!t) return e_fail(em, YAML_MEMORY_ERROR, "out of memory");
        e->buf = t; e->cap = nc;
    }
    e->buf[e->len++] = (unsigned char)c;
    e->column++;
    return 1;
}
static int w_break(yaml_emitter_t *em)
{
    if (!w_put(em, '\n')) return 0;
    em->impl->column = 0; em->impl->line++;
    return 1;
}
/* one UTF-8 character from s; only its first byte counts as a column (like libyaml's WRITE) */
static size_t u8width(unsigned char c)
{
    return (c & 0x80) == 0 ? 1 : (c & 0xE0) == 0xC0 ? 2 : (c & 0xF0) == 0xE0 ? 3 : (c & 0xF8) == 0xF0 ? 4 : 1;
}
static int w_char(yaml_emitter_t *em, const unsigned char *s, size_t w)
{
    int col = em->impl->column;
    for (size_t i = 0; i < w; ++i) if (!w_put(em, s[i])) return 0;
    em->impl->column = col + 1;
    return 1;
}

static int push_int(int **a, size_t *n, size_t *cap, int v)
{
    if (*n == *cap) {
        size_t nc = *cap ? *cap * 2 : 32;
        int *t = (int *)realloc(*a, nc * sizeof(int));
        if (!t) return 0;
        *a = t; *cap = nc;
    }
    (*a)[(*n)++] = v;
    return 1;
}
#define PUSH_STATE(em, s) (push_int(&(em)->impl->states, &(em)->impl->nstates, &(em)->impl->cstates, (s)) \
                           || e_fail((em), YAML_MEMORY_ERROR, "out of memory"))
static int pop_state(E *e) { return e->nstates ? e->states[--e->nstates] : ES_END; }
static int pop_indent(E *e) { return e->nindents ? e->indents[--e->nindents] : -1; }

static int increase_indent(yaml_emitter_t *em, int flow, int indentless)
{
    E *e = em->impl;
    if (!push_int(&e->indents, &e->nindents, &e->cindents, e->indent))
        return e_fail(em, YAML_MEMORY_ERROR, "out of memory");
    if (e->indent < 0) e->indent = flow ? e->best_indent : 0;
    else if (!indentless) e->indent += e->best_indent;
    return 1;
}

static int write_indent(yaml_emitter_t *em)
{
    E *e = em->impl;
    int indent = e->indent >= 0 ? e->indent : 0;
    if (!e->indention || e->column > indent || (e->column == indent && !e->whitespace))
        if (!w_break(em)) return 0;
    while (e->column < indent) if (!w_put(em, ' ')) return 0;
    e->whitespace = 1; e->indention = 1;
    return 1;
}

static int write_indicator(yaml_emitter_t *em, const char *ind, int need_whitespace, int is_whitespace, int is_indention)
{
    E *e = em->impl;
    if (need_whitespace && !e->whitespace) if (!w_put(em, ' ')) return 0;
    for (; *ind; ++ind) if (!w_put(em, *ind)) return 0;
    e->whitespace = is_whitespace;
    e->indention = (e->indention && is_indention);
    e->open_ended = 0;
    return 1;
}

static int c_printable(const unsigned char *s, size_t left)
{
    unsigned char a = s[0];
    if (a == 0x0A || (a >= 0x20 && a <= 0x7E)) return 1;
    if (left < 2) return 0;
    unsigned char b = s[1];
    if (a == 0xC2 && b >= 0xA0) return 1;
    if (a > 0xC2 && a < 0xED) return 1;
    if (a == 0xED && b < 0xA0) return 1;
    if (a == 0xEE) return 1;
    if (a == 0xEF && left >= 3 && !(b == 0xBB && s[2] == 0xBF) && !(b == 0xBF && (s[2] == 0xBE || s[2] == 0xBF))) return 1;
    return 0;
}
/* libyaml's IS_BREAK: CR, LF, NEL, LS, PS */
static int c_break(const unsigned char *s, size_t left)
{
    if (s[0] == '\r' || s[0] == '\n') return 1;
    if (left >= 2 && s[0] == 0xC2 && s[1] == 0x85) return 1;
    if (left >= 3 && s[0] == 0xE2 && s[1] == 0x80 && (s[2] == 0xA8 || s[2] == 0xA9)) return 1;
    return 0;
}
static int c_blankz(const unsigned char *s, size_t left)
{
    if (left == 0) return 1;
    return s[0] == ' ' || s[0] == '\t' || s[0] == 0 || c_break(s, left);
}

static void analyze_scalar(E *e, const unsigned char *v, size_t length)
{
    int block_indicators = 0, flow_indicators = 0, line_breaks = 0, special_characters = 0;
    int leading_space = 0, leading_break = 0, trailing_space = 0, trailing_break = 0;
    int break_space = 0, space_break = 0;
    int preceded_by_whitespace, followed_by_whitespace, previous_space = 0, previous_break = 0;
    e->sd.value = v; e->sd.length = length;
    if (!length) {
        e->sd.multiline = 0; e->sd.flow_plain_allowed = 0; e->sd.block_plain_allowed = 1;
        e->sd.single_quoted_allowed = 1; e->sd.block_allowed = 0;
        return;
    }
    if (length >= 3 && ((v[0] == '-' && v[1] == '-' && v[2] == '-') || (v[0] == '.' && v[1] == '.' && v[2] == '.'))) {
        block_indicators = 1; flow_indicators = 1;
    }
    preceded_by_whitespace = 1;
    {
        size_t w = u8width(v[0]);
        followed_by_whitespace = (w >= length) ? 1 : c_blankz(v + w, length - w);
    }
    size_t i = 0;
    while (i < length) {
        const unsigned char *s = v + i;
        size_t left = length - i, w = u8width(s[0]);
        if (w > left) w = left;
        int c = s[0];
        if (i == 0) {
            if (c == '#' || c == ',' || c == '[' || c == ']' || c == '{' || c == '}' || c == '&' || c == '*'
                || c == '!' || c == '|' || c == '>' || c == '\'' || c == '"' || c == '%' || c == '@' || c == '`') {
                flow_indicators = 1; block_indicators = 1;
            }
            if (c == '?' || c == ':') {
                flow_indicators = 1;
                if (followed_by_whitespace) block_indicators = 1;
            }
            if (c == '-' && followed_by_whitespace) { flow_indicators = 1; block_indicators = 1; }
        } else {
            if (c == ',' || c == '?' || c == '[' || c == ']' || c == '{' || c == '}') flow_indicators = 1;
            if (c == ':') {
                flow_indicators = 1;
                if (followed_by_whitespace) block_indicators = 1;
            }
            if (c == '#' && preceded_by_whitespace) { flow_indicators = 1; block_indicators = 1; }
        }
        if (!c_printable(s, left) || (c & 0x80)) special_characters = 1;   /* unicode output is off */
        if (c_break(s, left)) line_breaks = 1;
        if (c == ' ') {
            if (i == 0) leading_space = 1;
            if (i + w == length) trailing_space = 1;
            if (previous_break) break_space = 1;
            previous_space = 1; previous_break = 0;
        } else if (c_break(s, left)) {
            if (i == 0) leading_break = 1;
            if (i + w == length) trailing_break = 1;
            if (previous_space) space_break = 1;
            previous_space = 0; previous_break = 1;
        } else { previous_space = 0; previous_break = 0; }
        preceded_by_whitespace = c_blankz(s, left);
        i += w;
        if (i < length) {
            size_t w2 = u8width(v[i]);
            followed_by_whitespace = (i + w2 >= length) ? 1 : c_blankz(v + i + w2, length - i - w2);
        }
    }
    e->sd.multiline = line_breaks;
    e->sd.flow_plain_allowed = e->sd.block_plain_allowed = e->sd.single_quoted_allowed = e->sd.block_allowed = 1;
    if (leading_space || leading_break || trailing_space || trailing_break)
        e->sd.flow_plain_allowed = e->sd.block_plain_allowed = 0;
    if (trailing_space) e->sd.block_allowed = 0;
    if (break_space) e->sd.flow_plain_allowed = e->sd.block_plain_allowed = e->sd.single_quoted_allowed = 0;
    if (space_break || special_characters)
        e->sd.flow_plain_allowed = e->sd.block_plain_allowed = e->sd.single_quoted_allowed = e->sd.block_allowed = 0;
    if (line_breaks) e->sd.flow_plain_allowed = e->sd.block_plain_allowed = 0;
    if (flow_indicators) e->sd.flow_plain_allowed = 0;
    if (block_indicators) e->sd.block_plain_allowed = 0;
}

static int select_scalar_style(yaml_emitter_t *em, yaml_event_t *ev)
{
    E *e = em->impl;
    yaml_scalar_style_t style = ev->data.scalar.style;
    /* tags are only analysed when neither implicit flag is set (emitter.c yaml_emitter_analyze_event) */
    int tagged = ev->data.scalar.tag && !ev->data.scalar.plain_implicit && !ev->data.scalar.quoted_implicit;
    if (tagged || ev->data.scalar.anchor)
        return e_fail(em, YAML_EMITTER_ERROR, "anchors and explicit tags are not supported");
    if (!ev->data.scalar.plain_implicit && !ev->data.scalar.quoted_implicit)
        return e_fail(em, YAML_EMITTER_ERROR, "neither tag nor implicit flags are specified");
    if (style == YAML_ANY_SCALAR_STYLE) style = YAML_PLAIN_SCALAR_STYLE;
    if (e->simple_key_context && e->sd.multiline) style = YAML_DOUBLE_QUOTED_SCALAR_STYLE;
    if (style == YAML_PLAIN_SCALAR_STYLE) {
        if ((e->flow_level && !e->sd.flow_plain_allowed) || (!e->flow_level && !e->sd.block_plain_allowed))
            style = YAML_SINGLE_QUOTED_SCALAR_STYLE;
        if (!e->sd.length && (e->flow_level || e->simple_key_context)) style = YAML_SINGLE_QUOTED_SCALAR_STYLE;
        if (!ev->data.scalar.plain_implicit) style = YAML_SINGLE_QUOTED_SCALAR_STYLE;
    }
    if (style == YAML_SINGLE_QUOTED_SCALAR_STYLE && !e->sd.single_quoted_allowed)
        style = YAML_DOUBLE_QUOTED_SCALAR_STYLE;
    if (style == YAML_LITERAL_SCALAR_STYLE || style == YAML_FOLDED_SCALAR_STYLE)
        style = YAML_DOUBLE_QUOTED_SCALAR_STYLE;       /* block scalars: out of scope, the safe style instead */
    /* a scalar that may only be plain-implicit but cannot be written plain gets the '!' tag */
    e->sd.bang = (!ev->data.scalar.quoted_implicit && style != YAML_PLAIN_SCALAR_STYLE);
    e->sd.style = style;
    return 1;
}

static int write_plain(yaml_emitter_t *em, const unsigned char *v, size_t length, int allow_breaks)
{
    E *e = em->impl;
    int spaces = 0, breaks = 0;
    if (!e->whitespace && (length || e->flow_level)) if (!w_put(em, ' ')) return 0;
    for (size_t i = 0; i < length;) {
        size_t left = length - i, w = u8width(v[i]);
        if (w > left) w = left;
        if (v[i] == ' ') {
            if (allow_breaks && !spaces && e->column > e->best_width && !(i + 1 < length && v[i + 1] == ' ')) {
                if (!write_indent(em)) return 0;
            } else if (!w_put(em, ' ')) return 0;
            spaces = 1;
        } else if (c_break(v + i, left)) {
            if (!breaks && v[i] == '\n') if (!w_break(em)) return 0;
            if (!w_break(em)) return 0;
            e->indention = 1; breaks = 1;
        } else {
            if (breaks) if (!write_indent(em)) return 0;
            if (!w_char(em, v + i, w)) return 0;
            e->indention = 0; spaces = 0; breaks = 0;
        }
        i += w;
    }
    e->whitespace = 0; e->indention = 0;
    return 1;
}

static int write_single_quoted(yaml_emitter_t *em, const unsigned char *v, size_t length, int allow_breaks)
{
    E *e = em->impl;
    int spaces = 0, breaks = 0;
    if (!write_indicator(em, "'", 1, 0, 0)) return 0;
    for (size_t i = 0; i < length;) {
        size_t left = length - i, w = u8width(v[i]);
        if (w > left) w = left;
        if (v[i] == ' ') {
            if (allow_breaks && !spaces && e->column > e->best_width && i != 0 && i != length - 1
                && !(i + 1 < length && v[i + 1] == ' ')) {
                if (!write_indent(em)) return 0;
            } else if (!w_put(em, ' ')) return 0;
            spaces = 1;
        } else if (c_break(v + i, left)) {
            if (!breaks && v[i] == '\n') if (!w_break(em)) return 0;
            if (!w_break(em)) return 0;
            e->indention = 1; breaks = 1;
        } else {
            if (breaks) if (!write_indent(em)) return 0;
            if (v[i] == '\'') if (!w_put(em, '\'')) return 0;
            if (!w_char(em, v + i, w)) return 0;
            e->indention = 0; spaces = 0; breaks = 0;
        }
        i += w;
    }
    if (breaks) if (!write_indent(em)) return 0;
    if (!write_indicator(em, "'", 0, 0, 0)) return 0;
    e->whitespace = 0; e->indention = 0;
    return 1;
}

static int write_double_quoted(yaml_emitter_t *em, const unsigned char *v, size_t length, int allow_breaks)
{
    E *e = em->impl;
    int spaces = 0;
    if (!write_indicator(em, "\"", 1, 0, 0)) return 0;
    for (size_t i = 0; i < length;) {
        size_t left = length - i, w = u8width(v[i]);
        if (w > left) w = left;
        int c = v[i];
        if (!c_printable(v + i, left) || (c & 0x80) || c_break(v + i, left) || c == '"' || c == '\\') {
            unsigned value = (w == 1) ? (unsigned)c : (w == 2) ? (c & 0x1Fu) : (w == 3) ? (c & 0x0Fu) : (c & 0x07u);
            for (size_t k = 1; k < w; ++k) value = (value << 6) | (v[i + k] & 0x3Fu);
            if (!w_put(em, '\\')) return 0;
            int ch = 0;
            switch (value) {
            case 0x00: ch = '0'; break;  case 0x07: ch = 'a'; break;  case 0x08: ch = 'b'; break;
            case 0x09: ch = 't'; break;  case 0x0A: ch = 'n'; break;  case 0x0B: ch = 'v'; break;
            case 0x0C: ch = 'f'; break;  case 0x0D: ch = 'r'; break;  case 0x1B: ch = 'e'; break;
            case 0x22: ch = '"'; break;  case 0x5C: ch = '\\'; break; case 0x85: ch = 'N'; break;
            case 0xA0: ch = '_'; break;  case 0x2028: ch = 'L'; break; case 0x2029: ch = 'P'; break;
            default: break;
            }
            if (ch) { if (!w_put(em, ch)) return 0; }
            else {
                int width = value <= 0xFF ? 2 : value <= 0xFFFF ? 4 : 8;
                if (!w_put(em, width == 2 ? 'x' : width == 4 ? 'u' : 'U')) return 0;
                for (int k = (width - 1) * 4; k >= 0; k -= 4) {
                    unsigned d = (value >> k) & 0x0F;
                    if (!w_put(em, (int)(d + (d < 10 ? '0' : 'A' - 10)))) return 0;
                }
            }
            spaces = 0;
        } else if (c == ' ') {
            if (allow_breaks && !spaces && e->column > e->best_width && i != 0 && i != length - 1) {
                if (!write_indent(em)) return 0;
                if (i + 1 < length && v[i + 1] == ' ') if (!w_put(em, '\\')) return 0;
            } else if (!w_put(em, ' ')) return 0;
            spaces = 1;
        } else {
            if (!w_char(em, v + i, w)) return 0;
            spaces = 0;
        }
        i += w;
    }
    if (!write_indicator(em, "\"", 0, 0, 0)) return 0;
    e->whitespace = 0; e->indention = 0;
    return 1;
}

static int check_empty(E *e, yaml_event_type_t start, yaml_event_type_t end)
{
    if (e->qtail - e->qhead < 2) return 0;
    return e->q[e->qhead].type == start && e->q[e->qhead + 1].type == end;
}

static int check_simple_key(E *e)
{
    yaml_event_t *ev = &e->q[e->qhead];
    size_t length = 0;
    switch (ev->type) {
    case YAML_SCALAR_EVENT:
        if (e->sd.multiline) return 0;
        length += e->sd.length;
        break;
    case YAML_SEQUENCE_START_EVENT:
        if (!check_empty(e, YAML_SEQUENCE_START_EVENT, YAML_SEQUENCE_END_EVENT)) return 0;
        break;
    case YAML_MAPPING_START_EVENT:
        if (!check_empty(e, YAML_MAPPING_START_EVENT, YAML_MAPPING_END_EVENT)) return 0;
        break;
    default: return 0;
    }
    return length <= 128;
}

static int emit_node(yaml_emitter_t *em, yaml_event_t *ev, int root, int sequence, int mapping, int simple_key)
{
    E *e = em->impl;
    e->root_context = root; e->sequence_context = sequence;
    e->mapping_context = mapping; e->simple_key_context = simple_key;
    switch (ev->type) {
    case YAML_SCALAR_EVENT: {
        if (!select_scalar_style(em, ev)) return 0;
        if (e->sd.bang) {                       /* yaml_emitter_write_tag_handle("!") */
            if (!e->whitespace) if (!w_put(em, ' ')) return 0;
            if (!w_put(em, '!')) return 0;
            e->whitespace = 0; e->indention = 0;
        }
        if (!increase_indent(em, 1, 0)) return 0;
        int ok;
        if (e->sd.style == YAML_PLAIN_SCALAR_STYLE) ok = write_plain(em, e->sd.value, e->sd.length, !e->simple_key_context);
        else if (e->sd.style == YAML_SINGLE_QUOTED_SCALAR_STYLE) ok = write_single_quoted(em, e->sd.value, e->sd.length, !e->simple_key_context);
        else ok = write_double_quoted(em, e->sd.value, e->sd.length, !e->simple_key_context);
        if (!ok) return 0;
        e->indent = pop_indent(e);
        e->state = pop_state(e);
        return 1;
    }
    case YAML_SEQUENCE_START_EVENT:
        if (ev->data.sequence_start.anchor || (ev->data.sequence_start.tag && !ev->data.sequence_start.implicit))
            return e_fail(em, YAML_EMITTER_ERROR, "anchors and explicit tags are not supported");
        if (e->flow_level || ev->data.sequence_start.style == YAML_FLOW_SEQUENCE_STYLE
            || check_empty(e, YAML_SEQUENCE_START_EVENT, YAML_SEQUENCE_END_EVENT))
            e->state = ES_FLOW_SEQ_FIRST_ITEM;
        else e->state = ES_BLOCK_SEQ_FIRST_ITEM;
        return 1;
    case YAML_MAPPING_START_EVENT:
        if (ev->data.mapping_start.anchor || (ev->data.mapping_start.tag && !ev->data.mapping_start.implicit))
            return e_fail(em, YAML_EMITTER_ERROR, "anchors and explicit tags are not supported");
        if (e->flow_level || ev->data.mapping_start.style == YAML_FLOW_MAPPING_STYLE
            || check_empty(e, YAML_MAPPING_START_EVENT, YAML_MAPPING_END_EVENT))
            e->state = ES_FLOW_MAP_FIRST_KEY;
        else e->state = ES_BLOCK_MAP_FIRST_KEY;
        return 1;
    case YAML_ALIAS_EVENT:
        return e_fail(em, YAML_EMITTER_ERROR, "aliases are not supported");
    default:
        return e_fail(em, YAML_EMITTER_ERROR, "expected SCALAR, SEQUENCE-START, MAPPING-START, or ALIAS");
    }
}

static int state_machine(yaml_emitter_t *em, yaml_event_t *ev)
{
    E *e = em->impl;
    int first = 0;
    switch (e->state) {
    case ES_STREAM_START:
        if (ev->type != YAML_STREAM_START_EVENT) return e_fail(em, YAML_EMITTER_ERROR, "expected STREAM-START");
        if (e->best_indent < 2 || e->best_indent > 9) e->best_indent = 2;
        if (e->best_width >= 0 && e->best_width <= e->best_indent * 2) e->best_width = 80;
        if (e->best_width < 0) e->best_width = 0x7fffffff;
        e->indent = -1; e->line = 0; e->column = 0; e->whitespace = 1; e->indention = 1;
        e->state = ES_FIRST_DOCUMENT_START;
        return 1;

    case ES_FIRST_DOCUMENT_START: first = 1; /* fall through */
    case ES_DOCUMENT_START:
        if (ev->type == YAML_DOCUMENT_START_EVENT) {
            int implicit = ev->data.document_start.implicit;
            if (!first) implicit = 0;
            if (!implicit) {
                if (!write_indent(em)) return 0;
                if (!write_indicator(em, "---", 1, 0, 0)) return 0;
            }
            e->state = ES_DOCUMENT_CONTENT;
            e->open_ended = 0;
            return 1;
        }
        if (ev->type == YAML_STREAM_END_EVENT) {
            if (e->open_ended == 2) {
                if (!write_indicator(em, "...", 1, 0, 0)) return 0;
                e->open_ended = 0;
                if (!write_indent(em)) return 0;
            }
            if (!yaml_emitter_flush(em)) return 0;
            e->state = ES_END;
            return 1;
        }
        return e_fail(em, YAML_EMITTER_ERROR, "expected DOCUMENT-START or STREAM-END");

    case ES_DOCUMENT_CONTENT:
        if (!PUSH_STATE(em, ES_DOCUMENT_END)) return 0;
        return emit_node(em, ev, 1, 0, 0, 0);

    case ES_DOCUMENT_END:
        if (ev->type != YAML_DOCUMENT_END_EVENT) return e_fail(em, YAML_EMITTER_ERROR, "expected DOCUMENT-END");
        if (!write_indent(em)) return 0;
        if (!ev->data.document_end.implicit) {
            if (!write_indicator(em, "...", 1, 0, 0)) return 0;
            e->open_ended = 0;
            if (!write_indent(em)) return 0;
        } else if (!e->open_ended) e->open_ended = 1;
        if (!yaml_emitter_flush(em)) return 0;
        e->state = ES_DOCUMENT_START;
        return 1;

    case ES_FLOW_SEQ_FIRST_ITEM: first = 1; /* fall through */
    case ES_FLOW_SEQ_ITEM:
        if (first) {
            if (!write_indicator(em, "[", 1, 1, 0)) return 0;
            if (!increase_indent(em, 1, 0)) return 0;
            e->flow_level++;
        }
        if (ev->type == YAML_SEQUENCE_END_EVENT) {
            e->flow_level--;
            e->indent = pop_indent(e);
            if (!write_indicator(em, "]", 0, 0, 0)) return 0;
            e->state = pop_state(e);
            return 1;
        }
        if (!first) if (!write_indicator(em, ",", 0, 0, 0)) return 0;
        if (e->column > e->best_width) if (!write_indent(em)) return 0;
        if (!PUSH_STATE(em, ES_FLOW_SEQ_ITEM)) return 0;
        return emit_node(em, ev, 0, 1, 0, 0);

    case ES_FLOW_MAP_FIRST_KEY: first = 1; /* fall through */
    case ES_FLOW_MAP_KEY:
        if (first) {
            if (!write_indicator(em, "{", 1, 1, 0)) return 0;
            if (!increase_indent(em, 1, 0)) return 0;
            e->flow_level++;
        }
        if (ev->type == YAML_MAPPING_END_EVENT) {
            e->flow_level--;
            e->indent = pop_indent(e);
            if (!write_indicator(em, "}", 0, 0, 0)) return 0;
            e->state = pop_state(e);
            return 1;
        }
        if (!first) if (!write_indicator(em, ",", 0, 0, 0)) return 0;
        if (e->column > e->best_width) if (!write_indent(em)) return 0;
        if (check_simple_key(e)) {
            if (!PUSH_STATE(em, ES_FLOW_MAP_SIMPLE_VALUE)) return 0;
            return emit_node(em, ev, 0, 0, 1, 1);
        }
        if (!write_indicator(em, "?", 1, 0, 0)) return 0;
        if (!PUSH_STATE(em, ES_FLOW_MAP_VALUE)) return 0;
        return emit_node(em, ev, 0, 0, 1, 0);

    case ES_FLOW_MAP_SIMPLE_VALUE:
        if (!write_indicator(em, ":", 0, 0, 0)) return 0;
        if (!PUSH_STATE(em, ES_FLOW_MAP_KEY)) return 0;
        return emit_node(em, ev, 0, 0, 1, 0);
    case ES_FLOW_MAP_VALUE:
        if (e->column > e->best_width) if (!write_indent(em)) return 0;
        if (!write_indicator(em, ":", 1, 0, 0)) return 0;
        if (!PUSH_STATE(em, ES_FLOW_MAP_KEY)) return 0;
        return emit_node(em, ev, 0, 0, 1, 0);

    case ES_BLOCK_SEQ_FIRST_ITEM: first = 1; /* fall through */
    case ES_BLOCK_SEQ_ITEM:
        if (first) if (!increase_indent(em, 0, (e->mapping_context && !e->indention))) return 0;
        if (ev->type == YAML_SEQUENCE_END_EVENT) {
            e->indent = pop_indent(e);
            e->state = pop_state(e);
            return 1;
        }
        if (!write_indent(em)) return 0;
        if (!write_indicator(em, "-", 1, 0, 1)) return 0;
        if (!PUSH_STATE(em, ES_BLOCK_SEQ_ITEM)) return 0;
        return emit_node(em, ev, 0, 1, 0, 0);

    case ES_BLOCK_MAP_FIRST_KEY: first = 1; /* fall through */
    case ES_BLOCK_MAP_KEY:
        if (first) if (!increase_indent(em, 0, 0)) return 0;
        if (ev->type == YAML_MAPPING_END_EVENT) {
            e->indent = pop_indent(e);
            e->state = pop_state(e);
            return 1;
        }
        if (!write_indent(em)) return 0;
        if (check_simple_key(e)) {
            if (!PUSH_STATE(em, ES_BLOCK_MAP_SIMPLE_VALUE)) return 0;
            return emit_node(em, ev, 0, 0, 1, 1);
        }
        if (!write_indicator(em, "?", 1, 0, 1)) return 0;
        if (!PUSH_STATE(em, ES_BLOCK_MAP_VALUE)) return 0;
        return emit_node(em, ev, 0, 0, 1, 0);

    case ES_BLOCK_MAP_SIMPLE_VALUE:
        if (!write_indicator(em, ":", 0, 0, 0)) return 0;
        if (!PUSH_STATE(em, ES_BLOCK_MAP_KEY)) return 0;
        return emit_node(em, ev, 0, 0, 1, 0);
    case ES_BLOCK_MAP_VALUE:
        if (!write_indent(em)) return 0;
        if (!write_indicator(em, ":", 1, 0, 1)) return 0;
        if (!PUSH_STATE(em, ES_BLOCK_MAP_KEY)) return 0;
        return emit_node(em, ev, 0, 0, 1, 0);

    case ES_END:
    default:
        return e_fail(em, YAML_EMITTER_ERROR, "expected nothing after STREAM-END");
    }
}

static int need_more_events(E *e)
{
    if (e->qhead == e->qtail) return 1;
    size_t accumulate;
    switch (e->q[e->qhead].type) {
    case YAML_DOCUMENT_START_EVENT: accumulate = 1; break;
    case YAML_SEQUENCE_START_EVENT: accumulate = 2; break;
    case YAML_MAPPING_START_EVENT: accumulate = 3; break;
    default: return 0;
    }
    if (e->qtail - e->qhead > accumulate) return 0;
    int level = 0;
    for (size_t i = e->qhead; i < e->qtail; ++i) {
        switch (e->q[i].type) {
        case YAML_STREAM_START_EVENT: case YAML_DOCUMENT_START_EVENT:
        case YAML_SEQUENCE_START_EVENT: case YAML_MAPPING_START_EVENT: level++; break;
        case YAML_STREAM_END_EVENT: case YAML_DOCUMENT_END_EVENT:
        case YAML_SEQUENCE_END_EVENT: case YAML_MAPPING_END_EVENT: level--; break;
        default: break;
        }
        if (!level) return 0;
    }
    return 1;
}

int yaml_emitter_emit(yaml_emitter_t *em, yaml_event_t *event)
{
    E *e = em->impl;
    if (!e) { yaml_event_delete(event); return 0; }
    if (e->qtail == e->qcap) {
        if (e->qhead > 0) {
            memmove(e->q, e->q + e->qhead, (e->qtail - e->qhead) * sizeof(yaml_event_t));
            e->qtail -= e->qhead; e->qhead = 0;
        } else {
            size_t nc = e->qcap ? e->qcap * 2 : 16;
            yaml_event_t *t = (yaml_event_t *)realloc(e->q, nc * sizeof(yaml_event_t));
            if (!t) { yaml_event_delete(event); return e_fail(em, YAML_MEMORY_ERROR, "out of memory"); }
            e->q = t; e->qcap = nc;
        }
    }
    e->q[e->qtail++] = *event;
    memset(event, 0, sizeof(*event));
    while (!need_more_events(e)) {
        yaml_event_t *head = &e->q[e->qhead];
        if (head->type == YAML_SCALAR_EVENT)
            analyze_scalar(e, head->data.scalar.value, head->data.scalar.length);
        int ok = state_machine(em, head);
        yaml_event_delete(head);
        e->qhead++;
        if (!ok) return 0;
    }
    return 1;
}

/* -------------------------------------------------- event listings (tests) */

typedef struct { char *s; size_t n, cap; int oom; } sbuf;
static void sb_put(sbuf *b, const char *s, size_t n)
{
    if (b->oom) return;
    if (b->n + n + 1 > b->cap) {
        size_t nc = b->cap ? b->cap : 1024;
        while (nc < b->n + n + 1) nc *= 2;
        char *t = (char *)realloc(b->s, nc);
        if (!t) { b->oom = 1; return; }
        b->s = t; b->cap = nc;
    }
    memcpy(b->s + b->n, s, n);
    b->n += n;
    b->s[b->n] = 0;
}
static void sb_str(sbuf *b, const char *s) { sb_put(b, s, strlen(s)); }

int ylite_events_from_yaml(const unsigned char *input, size_t size, char **listing, size_t *listing_size)
{
    yaml_parser_t parser;
    yaml_event_t ev;
    sbuf b = {0};
    int rc = 0;
    if (!yaml_parser_initialize(&parser)) return YAML_MEMORY_ERROR;
    yaml_parser_set_input_string(&parser, input, size);
    sb_str(&b, "");
    for (;;) {
        if (!yaml_parser_parse(&parser, &ev)) {
            char tmp[256];
            snprintf(tmp, sizeof(tmp), "!ERR %d %zu %s\n", (int)parser.error, parser.problem_mark.line,
                     parser.problem ? parser.problem : "");
            sb_str(&b, tmp);
            rc = parser.error;
            break;
        }
        yaml_event_type_t t = ev.type;
        switch (t) {
        case YAML_STREAM_START_EVENT: sb_str(&b, "+STR\n"); break;
        case YAML_STREAM_END_EVENT: sb_str(&b, "-STR\n"); break;
        case YAML_DOCUMENT_START_EVENT: sb_str(&b, ev.data.document_start.implicit ? "+DOC\n" : "+DOC ---\n"); break;
        case YAML_DOCUMENT_END_EVENT: sb_str(&b, ev.data.document_end.implicit ? "-DOC\n" : "-DOC ...\n"); break;
        case YAML_MAPPING_START_EVENT:
            sb_str(&b, ev.data.mapping_start.style == YAML_FLOW_MAPPING_STYLE ? "+MAP {}\n" : "+MAP\n"); break;
        case YAML_MAPPING_END_EVENT: sb_str(&b, "-MAP\n"); break;
        case YAML_SEQUENCE_START_EVENT:
            sb_str(&b, ev.data.sequence_start.style == YAML_FLOW_SEQUENCE_STYLE ? "+SEQ []\n" : "+SEQ\n"); break;
        case YAML_SEQUENCE_END_EVENT: sb_str(&b, "-SEQ\n"); break;
        case YAML_SCALAR_EVENT: {
            char head[16];
            static const char sty[] = "apsdlf";
            snprintf(head, sizeof(head), "=VAL %d%d%c ", ev.data.scalar.plain_implicit, ev.data.scalar.quoted_implicit,
                     sty[ev.data.scalar.style]);
            sb_str(&b, head);
            for (size_t i = 0; i < ev.data.scalar.length; ++i) {
                unsigned char c = ev.data.scalar.value[i];
                if (c == '\\') sb_str(&b, "\\\\");
                else if (c == '\n') sb_str(&b, "\\n");
                else if (c == '\r') sb_str(&b, "\\r");
                else if (c == '\t') sb_str(&b, "\\t");
                else if (c == 0) sb_str(&b, "\\0");
                else sb_put(&b, (const char *)&c, 1);
            }
            sb_str(&b, "\n");
            break;
        }
        default: break;
        }
        yaml_event_delete(&ev);
        if (t == YAML_STREAM_END_EVENT || t == YAML_NO_EVENT) break;
    }
    yaml_parser_delete(&parser);
    if (b.oom) { free(b.s); return YAML_MEMORY_ERROR; }
    *listing = b.s; *listing_size = b.n;
    return rc;
}

struct sink { sbuf b; };
static int sink_write(void *data, unsigned char *buffer, size_t size)
{
    struct sink *s = (struct sink *)data;
    sb_put(&s->b, (const char *)buffer, size);
    return !s->b.oom;
}

int ylite_yaml_from_events(const char *listing, size_t size, int width, char **text, size_t *text_size)
{
    yaml_emitter_t em;
    yaml_event_t ev;
    struct sink sink;
    memset(&sink, 0, sizeof(sink));
    if (!yaml_emitter_initialize(&em)) return YAML_MEMORY_ERROR;
    yaml_emitter_set_output(&em, sink_write, &sink);
    if (width) yaml_emitter_set_width(&em, width);
    sb_str(&sink.b, "");
    int rc = 0;
    size_t i = 0;
    unsigned char *val = (unsigned char *)malloc(size + 1);
    if (!val) { yaml_emitter_delete(&em); return YAML_MEMORY_ERROR; }
    while (i < size && !rc) {
        size_t j = i;
        while (j < size && listing[j] != '\n') j++;
        const char *l = listing + i;
        size_t n = j - i;
        i = j + 1;
        if (n < 4) continue;
        int ok = 1, have = 1;
        if (!strncmp(l, "+STR", 4)) yaml_stream_start_event_initialize(&ev, YAML_UTF8_ENCODING);
        else if (!strncmp(l, "-STR", 4)) yaml_stream_end_event_initialize(&ev);
        else if (!strncmp(l, "+DOC", 4)) yaml_document_start_event_initialize(&ev, NULL, NULL, NULL, !(n >= 8 && !strncmp(l + 4, " ---", 4)));
        else if (!strncmp(l, "-DOC", 4)) yaml_document_end_event_initialize(&ev, !(n >= 8 && !strncmp(l + 4, " ...", 4)));
        else if (!strncmp(l, "+MAP", 4)) yaml_mapping_start_event_initialize(&ev, NULL, NULL, 1,
                    (n >= 7 && !strncmp(l + 4, " {}", 3)) ? YAML_FLOW_MAPPING_STYLE : YAML_ANY_MAPPING_STYLE);
        else if (!strncmp(l, "-MAP", 4)) yaml_mapping_end_event_initialize(&ev);
        else if (!strncmp(l, "+SEQ", 4)) yaml_sequence_start_event_initialize(&ev, NULL, NULL, 1,
                    (n >= 7 && !strncmp(l + 4, " []", 3)) ? YAML_FLOW_SEQUENCE_STYLE : YAML_ANY_SEQUENCE_STYLE);
        else if (!strncmp(l, "-SEQ", 4)) yaml_sequence_end_event_initialize(&ev);
        else if (!strncmp(l, "=VAL ", 5) && n >= 9) {
            int p = l[5] == '1', q = l[6] == '1';
            yaml_scalar_style_t st = l[7] == 'p' ? YAML_PLAIN_SCALAR_STYLE : l[7] == 's' ? YAML_SINGLE_QUOTED_SCALAR_STYLE
                                   : l[7] == 'd' ? YAML_DOUBLE_QUOTED_SCALAR_STYLE : YAML_ANY_SCALAR_STYLE;
            size_t m = 0;
            for (size_t k = 9; k < n; ++k) {
                if (l[k] == '\\' && k + 1 < n) {
                    char c = l[++k];
                    val[m++] = (unsigned char)(c == 'n' ? '\n' : c == 'r' ? '\r' : c == 't' ? '\t' : c == '0' ? 0 : c);
                } else val[m++] = (unsigned char)l[k];
            }
            ok = yaml_scalar_event_initialize(&ev, NULL, NULL, val, (int)m, p, q, st);
        } else have = 0;
        if (!have) continue;
        if (!ok) { rc = YAML_MEMORY_ERROR; break; }
        if (!yaml_emitter_emit(&em, &ev)) rc = em.error ? (int)em.error : YAML_EMITTER_ERROR;
    }
    free(val);
    yaml_emitter_delete(&em);
    if (sink.b.oom) { free(sink.b.s); return YAML_MEMORY_ERROR; }
    *text = sink.b.s; *text_size = sink.b.n;
    return rc;
}
