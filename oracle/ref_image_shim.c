/*
 * ref_image_shim.c -- TEST INFRASTRUCTURE (oracle side), not product code.
 *
 * The reference's image.c cannot be compiled in this image (it includes
 * png.h / jpeglib.h unconditionally, image.c:11-16).  The render path only
 * needs a dozen symbols from it (nm of ndt.o), none of which touch a codec
 * except image_save*.  This file provides those symbols with in-memory
 * behaviour so that the UNMODIFIED ndt.c render_image() (ndt.c:900) can run
 * and its fp64 framebuffer can be captured by the harness.
 *
 * Layout contract (image.h:81-89): row-major, pixel (x,y) at
 * pixels + (width*y + x)*pixel_width; fp64 RGBA pixels are 32 bytes
 * (image.h:22-26), u8 RGBA are 4 bytes (image.h:14-19).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <sys/time.h>
#include "matrix.h"
#include "image.h"

/* last framebuffer handed to image_save / image_save_bg */
static unsigned char *cap_px = NULL;
static int cap_w = 0, cap_h = 0, cap_pw = 0;

int ref_shim_captured(const unsigned char **px, int *w, int *h, int *pw)
{
    *px = cap_px; *w = cap_w; *h = cap_h; *pw = cap_pw;
    return cap_px != NULL;
}

void ref_shim_drop(void)
{
    free(cap_px);
    cap_px = NULL;
    cap_w = cap_h = cap_pw = 0;
}

static void capture(image_t *im)
{
    size_t bytes = (size_t)im->width * im->height * im->pixel_width;
    ref_shim_drop();
    cap_px = malloc(bytes ? bytes : 1);
    memcpy(cap_px, im->pixels, bytes);
    cap_w = im->width; cap_h = im->height; cap_pw = im->pixel_width;
}

static long px_off(image_t *im, int x, int y)
{
    if (x < 0 || y < 0 || x >= im->width || y >= im->height)
        return -1;
    return ((long)im->width * y + x) * im->pixel_width;
}

int image_init(image_t *im)
{
    memset(im, 0, sizeof *im);
    im->pixel_width = (int)sizeof(pixel_t);
    return 0;
}

int dbl_image_init(image_t *im)
{
    memset(im, 0, sizeof *im);
    im->pixel_width = (int)sizeof(dbl_pixel_t);
    return 0;
}

int image_set_size(image_t *im, int w, int h)
{
    free(im->pixels);
    im->pixels = calloc((size_t)w * h, im->pixel_width);
    im->allocated = w * h * im->pixel_width;
    im->width = w;
    im->height = h;
    return 0;
}

int image_free(image_t *im)
{
    free(im->pixels);
    memset(im, 0, sizeof *im);
    return 0;
}

int image_copy(image_t *dst, image_t *src)
{
    if (src->pixel_width == (int)sizeof(dbl_pixel_t))
        dbl_image_init(dst);
    else
        image_init(dst);
    image_set_size(dst, src->width, src->height);
    memcpy(dst->pixels, src->pixels, (size_t)src->allocated);
    return 0;
}

int dbl_image_set_pixel(image_t *im, int x, int y, dbl_pixel_t *c)
{
    long o = px_off(im, x, y);
    if (o < 0)
        return -1;
    if (im->pixel_width == (int)sizeof(dbl_pixel_t)) {
        memcpy(im->pixels + o, c, sizeof *c);
    } else {
        pixel_t q;
        pixel_d2c(q, *c);
        memcpy(im->pixels + o, &q, sizeof q);
    }
    return 0;
}

int dbl_image_get_pixel(image_t *im, int x, int y, dbl_pixel_t *c)
{
    long o = px_off(im, x, y);
    if (o < 0) {
        memset(c, 0, sizeof *c);
        return -1;
    }
    if (im->pixel_width == (int)sizeof(dbl_pixel_t)) {
        memcpy(c, im->pixels + o, sizeof *c);
    } else {
        pixel_t q;
        memcpy(&q, im->pixels + o, sizeof q);
        pixel_c2d(*c, q);
    }
    return 0;
}

int image_save(image_t *im, char *fname, int fmt)
{
    (void)fname; (void)fmt;
    capture(im);
    return 0;
}

int image_save_bg(image_t *im, char *fname, int fmt)
{
    return image_save(im, fname, fmt);
}

int image_active_saves(void)
{
    return 0;
}

int image_avg_dbl_pixels4(dbl_pixel_t *a, dbl_pixel_t *b, dbl_pixel_t *c,
                          dbl_pixel_t *d, dbl_pixel_t *avg, double *var)
{
    const double *p[4] = { &a->r, &b->r, &c->r, &d->r };
    double *o = &avg->r;
    double spread = 0.0;
    for (int ch = 0; ch < 4; ++ch) {
        o[ch] = (p[0][ch] + p[1][ch] + p[2][ch] + p[3][ch]) / 4;
        double s = 0.0;
        s += fabs(o[ch] - p[0][ch]) + fabs(o[ch] - p[1][ch]) +
             fabs(o[ch] - p[2][ch]) + fabs(o[ch] - p[3][ch]);
        spread += s;
    }
    if (var)
        *var = spread;
    return 0;
}

int dbl_image_normalize(image_t *norm, image_t *src)
{
    /* depth-map post-processing: not on the measured path; plain copy */
    return image_copy(norm, src);
}
