/* yaml_lite/yaml.h -- the part of libyaml's C API that ndt uses, without libyaml.
 *
 * ndt reads and writes scenes as YAML through libyaml's EVENT interface
 * (reference scene.c:573-1094 emitter, scene.c:1096-2175 pull parser; the MPI
 * layer ships scenes as the same text, ndt.c:1153-1246).  libyaml is an
 * optional, unpinned system package there (CMakeLists.txt:30) and absent from
 * this image, so YAML scenes (BASELINE config 5, scenes/yaml.c) could not be
 * loaded at all.  This header + ndt_b200/csrc/yaml_lite.c provide the same
 * type, enum and function names, so that the reference's scene.c and
 * scenes/yaml.c compile UNMODIFIED with `-DWITH_YAML -Iinclude/yaml_lite` and
 * make the very same object_add_* / scene_alloc_light calls for a given file.
 *
 * What it is: an event producer for the YAML subset scenes are written in
 * (block and flow mappings / sequences, complex "? " keys, plain scalars incl.
 * multi-line ones, quoted scalars, comments, multi-document streams) and an event consumer that writes the text libyaml
 * 0.2.5 writes for the same events (block/flow layout, 80-column folding of
 * flow sequences, scalar style selection and quoting).  Both are pinned against
 * libyaml 0.2.5 itself (PyYAML's CParser / CEmitter) in tests/test_yaml_lite.py.
 * What it is not: anchors, aliases, tags, directives, block scalars and flow
 * collections used as simple keys ("[a]: b") are refused with a parser error,
 * never guessed at.
 *
 * Link names carry a ylite_ prefix (macros below) so the library can share a
 * process with a real libyaml.
 */
#ifndef YAML_H
#define YAML_H

#include <stdio.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char yaml_char_t;

typedef struct yaml_version_directive_s { int major, minor; } yaml_version_directive_t;
typedef struct yaml_tag_directive_s { yaml_char_t *handle, *prefix; } yaml_tag_directive_t;

typedef enum yaml_encoding_e {
    YAML_ANY_ENCODING, YAML_UTF8_ENCODING, YAML_UTF16LE_ENCODING, YAML_UTF16BE_ENCODING
} yaml_encoding_t;

typedef enum yaml_break_e { YAML_ANY_BREAK, YAML_CR_BREAK, YAML_LN_BREAK, YAML_CRLN_BREAK } yaml_break_t;

typedef enum yaml_error_type_e {
    YAML_NO_ERROR, YAML_MEMORY_ERROR, YAML_READER_ERROR, YAML_SCANNER_ERROR,
    YAML_PARSER_ERROR, YAML_COMPOSER_ERROR, YAML_WRITER_ERROR, YAML_EMITTER_ERROR
} yaml_error_type_t;

typedef struct yaml_mark_s { size_t index, line, column; } yaml_mark_t;

typedef enum yaml_scalar_style_e {
    YAML_ANY_SCALAR_STYLE, YAML_PLAIN_SCALAR_STYLE, YAML_SINGLE_QUOTED_SCALAR_STYLE,
    YAML_DOUBLE_QUOTED_SCALAR_STYLE, YAML_LITERAL_SCALAR_STYLE, YAML_FOLDED_SCALAR_STYLE
} yaml_scalar_style_t;

typedef enum yaml_sequence_style_e {
    YAML_ANY_SEQUENCE_STYLE, YAML_BLOCK_SEQUENCE_STYLE, YAML_FLOW_SEQUENCE_STYLE
} yaml_sequence_style_t;

typedef enum yaml_mapping_style_e {
    YAML_ANY_MAPPING_STYLE, YAML_BLOCK_MAPPING_STYLE, YAML_FLOW_MAPPING_STYLE
} yaml_mapping_style_t;

typedef enum yaml_event_type_e {
    YAML_NO_EVENT,
    YAML_STREAM_START_EVENT, YAML_STREAM_END_EVENT,
    YAML_DOCUMENT_START_EVENT, YAML_DOCUMENT_END_EVENT,
    YAML_ALIAS_EVENT, YAML_SCALAR_EVENT,
    YAML_SEQUENCE_START_EVENT, YAML_SEQUENCE_END_EVENT,
    YAML_MAPPING_START_EVENT, YAML_MAPPING_END_EVENT
} yaml_event_type_t;

/* Same members as libyaml's yaml_event_t (scene.c reads .type and .data.scalar.value). */
typedef struct yaml_event_s {
    yaml_event_type_t type;
    union {
        struct { yaml_encoding_t encoding; } stream_start;
        struct {
            yaml_version_directive_t *version_directive;
            struct { yaml_tag_directive_t *start, *end; } tag_directives;
            int implicit;
        } document_start;
        struct { int implicit; } document_end;
        struct { yaml_char_t *anchor; } alias;
        struct {
            yaml_char_t *anchor, *tag, *value;
            size_t length;
            int plain_implicit, quoted_implicit;
            yaml_scalar_style_t style;
        } scalar;
        struct { yaml_char_t *anchor, *tag; int implicit; yaml_sequence_style_t style; } sequence_start;
        struct { yaml_char_t *anchor, *tag; int implicit; yaml_mapping_style_t style; } mapping_start;
    } data;
    yaml_mark_t start_mark, end_mark;
} yaml_event_t;

typedef int yaml_read_handler_t(void *data, unsigned char *buffer, size_t size, size_t *size_read);
typedef int yaml_write_handler_t(void *data, unsigned char *buffer, size_t size);

/* scene.c reads parser->error (scene.c:1105) and emitter->problem (scene.c:578). */
typedef struct yaml_parser_s {
    yaml_error_type_t error;
    const char *problem;
    size_t problem_offset;
    int problem_value;
    yaml_mark_t problem_mark;
    const char *context;
    yaml_mark_t context_mark;
    struct ylite_parser_impl *impl;
} yaml_parser_t;

typedef struct yaml_emitter_s {
    yaml_error_type_t error;
    const char *problem;
    struct ylite_emitter_impl *impl;
} yaml_emitter_t;

#define yaml_get_version_string              ylite_get_version_string
#define yaml_event_delete                    ylite_event_delete
#define yaml_stream_start_event_initialize   ylite_stream_start_event_initialize
#define yaml_stream_end_event_initialize     ylite_stream_end_event_initialize
#define yaml_document_start_event_initialize ylite_document_start_event_initialize
#define yaml_document_end_event_initialize   ylite_document_end_event_initialize
#define yaml_scalar_event_initialize         ylite_scalar_event_initialize
#define yaml_sequence_start_event_initialize ylite_sequence_start_event_initialize
#define yaml_sequence_end_event_initialize   ylite_sequence_end_event_initialize
#define yaml_mapping_start_event_initialize  ylite_mapping_start_event_initialize
#define yaml_mapping_end_event_initialize    ylite_mapping_end_event_initialize
#define yaml_parser_initialize               ylite_parser_initialize
#define yaml_parser_delete                   ylite_parser_delete
#define yaml_parser_set_input_string         ylite_parser_set_input_string
#define yaml_parser_set_input_file           ylite_parser_set_input_file
#define yaml_parser_parse                    ylite_parser_parse
#define yaml_emitter_initialize              ylite_emitter_initialize
#define yaml_emitter_delete                  ylite_emitter_delete
#define yaml_emitter_set_output_string       ylite_emitter_set_output_string
#define yaml_emitter_set_output_file         ylite_emitter_set_output_file
#define yaml_emitter_set_output              ylite_emitter_set_output
#define yaml_emitter_set_width               ylite_emitter_set_width
#define yaml_emitter_set_indent              ylite_emitter_set_indent
#define yaml_emitter_emit                    ylite_emitter_emit
#define yaml_emitter_flush                   ylite_emitter_flush

const char *yaml_get_version_string(void);

void yaml_event_delete(yaml_event_t *event);
int yaml_stream_start_event_initialize(yaml_event_t *event, yaml_encoding_t encoding);
int yaml_stream_end_event_initialize(yaml_event_t *event);
int yaml_document_start_event_initialize(yaml_event_t *event, yaml_version_directive_t *version_directive,
        yaml_tag_directive_t *tag_directives_start, yaml_tag_directive_t *tag_directives_end, int implicit);
int yaml_document_end_event_initialize(yaml_event_t *event, int implicit);
int yaml_scalar_event_initialize(yaml_event_t *event, const yaml_char_t *anchor, const yaml_char_t *tag,
        const yaml_char_t *value, int length, int plain_implicit, int quoted_implicit, yaml_scalar_style_t style);
int yaml_sequence_start_event_initialize(yaml_event_t *event, const yaml_char_t *anchor, const yaml_char_t *tag,
        int implicit, yaml_sequence_style_t style);
int yaml_sequence_end_event_initialize(yaml_event_t *event);
int yaml_mapping_start_event_initialize(yaml_event_t *event, const yaml_char_t *anchor, const yaml_char_t *tag,
        int implicit, yaml_mapping_style_t style);
int yaml_mapping_end_event_initialize(yaml_event_t *event);

int yaml_parser_initialize(yaml_parser_t *parser);
void yaml_parser_delete(yaml_parser_t *parser);
void yaml_parser_set_input_string(yaml_parser_t *parser, const unsigned char *input, size_t size);
void yaml_parser_set_input_file(yaml_parser_t *parser, FILE *file);
int yaml_parser_parse(yaml_parser_t *parser, yaml_event_t *event);

int yaml_emitter_initialize(yaml_emitter_t *emitter);
void yaml_emitter_delete(yaml_emitter_t *emitter);
void yaml_emitter_set_output_string(yaml_emitter_t *emitter, unsigned char *output, size_t size, size_t *size_written);
void yaml_emitter_set_output_file(yaml_emitter_t *emitter, FILE *file);
void yaml_emitter_set_output(yaml_emitter_t *emitter, yaml_write_handler_t *handler, void *data);
void yaml_emitter_set_width(yaml_emitter_t *emitter, int width);
void yaml_emitter_set_indent(yaml_emitter_t *emitter, int indent);
int yaml_emitter_emit(yaml_emitter_t *emitter, yaml_event_t *event);
int yaml_emitter_flush(yaml_emitter_t *emitter);

/* -- not in libyaml: an event-stream text form for tests and tools ------------------
 * One line per event, in the notation of the YAML test suite with the flags the
 * emitter needs made explicit:
 *   +STR | -STR | +DOC [---] | -DOC [...] | +MAP [{}] | -MAP | +SEQ [[]] | -SEQ
 *   =VAL <p><q><s> <text>    p, q = plain_implicit, quoted_implicit (0/1);
 *                            s = a(ny) p(lain) s(ingle) d(ouble); text with \\ \n \r \t \0 escapes
 * ylite_events_from_yaml: parse `input`, return the listing (malloc'd, caller frees with
 * ylite_free); a parse error ends the listing with "!ERR <code> <line> <problem>".
 * ylite_yaml_from_events: feed a listing to the emitter, return the text it writes.
 * Both return 0 on success, the yaml_error_type_t otherwise. */
int ylite_events_from_yaml(const unsigned char *input, size_t size, char **listing, size_t *listing_size);
int ylite_yaml_from_events(const char *listing, size_t size, int width, char **text, size_t *text_size);
void ylite_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* YAML_H */
