/* zng_ctx.c -- one zng_b200_ctx per host thread, created on first use.
 *
 * The reference keeps no global mutable state besides its functable (functable.c:17-39) and asks
 * for "one zng_stream per thread"; the GPU context (scratch buffers, streams) follows the same rule:
 * each host thread that calls into the library gets its own context on the thread's current CUDA
 * device, destroyed when the thread exits.
 */
#include "zng_b200.h"
#include "zng_host.h"
#include <pthread.h>
#include <stdlib.h>

static pthread_key_t ctx_key;
static pthread_once_t ctx_once = PTHREAD_ONCE_INIT;

static void ctx_destroy(void *p) { if (p) zng_b200_ctx_destroy((zng_b200_ctx *)p); }
static void ctx_key_init(void) { pthread_key_create(&ctx_key, ctx_destroy); }

zng_b200_ctx *zng_b200_thread_ctx(void) {
    pthread_once(&ctx_once, ctx_key_init);
    zng_b200_ctx *c = (zng_b200_ctx *)pthread_getspecific(ctx_key);
    if (c) return c;
    if (zng_b200_ctx_create(&c, -1) != ZNG_B200_OK) return NULL;   /* no device: the caller reports the error */
    pthread_setspecific(ctx_key, c);
    return c;
}
