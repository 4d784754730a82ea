/*
 * cloud_main.cpp — drop-in for the Cloud node's ./cloud executable (Cloud/cloud.c main(),
 * invoked with no arguments from cwd by Cloud/dragonfly_cipher_cloud.py:1233): reads cloud.key,
 * nbit.key, cloud.data and operator.txt, writes answer.data, exits 0 or 126.
 */
#include <cstdio>
#include <cstdlib>

#include "ieache_b200.h"

int main(int argc, char **argv)
{
    const char *dir = argc > 1 ? argv[1] : ".";
    const char *dev = getenv("IEACHE_DEVICE");
    ieache_ctx *ctx = nullptr;
    if (ieache_ctx_create(dev ? atoi(dev) : 0, &ctx) != IEACHE_OK) {
        fprintf(stderr, "cloud: %s\n", ieache_last_error());
        return 1;
    }
    printf("Reading the key...\n");
    double secs = 0;
    const int rc = ieache_cloud_run(ctx, dir, &secs);
    if (rc < 0) { fprintf(stderr, "cloud: %s\n", ieache_last_error()); ieache_ctx_destroy(ctx); return 1; }
    if (rc == 126) printf("Cannot multiply 256 bit number!\n");
    else printf("writing the answer to file...\n");
    ieache_ctx_destroy(ctx);
    return rc;
}
