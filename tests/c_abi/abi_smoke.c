/* abi_smoke.c — the C ABI of include/raytrace_b200.h driven from plain C99 (what a JNA / Panama / cgo binding sees):
 * a two-sphere world (a Lambertian ball on a large Lambertian ground, sky light), one traced ray with a closed-form
 * answer, one small render.  Exit codes: 0 ok, 3 no CUDA device (the library refuses: there is no CPU fallback),
 * 1 anything else.  Built and run by tests/test_c_abi.py.                                                          */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "raytrace_b200.h"

/* The JVM binding (clojure/native.clj) fills rt_scene_desc / rt_scene_ext / rt_path_bounce by BYTE OFFSET (JNA Memory
 * setInt / setPointer).  These literals are the ones native.clj hard-codes: if a field moves, this file stops compiling. */
#define OFF(type, field, at) _Static_assert(offsetof(type, field) == (at), #type "." #field " moved: update clojure/native.clj")
_Static_assert(sizeof(rt_scene_desc) == 112, "rt_scene_desc size: update clojure/native.clj");
OFF(rt_scene_desc, n_spheres, 0);    OFF(rt_scene_desc, center0_r, 8);     OFF(rt_scene_desc, center1, 16);
OFF(rt_scene_desc, t0t1, 24);        OFF(rt_scene_desc, sphere_flags, 32); OFF(rt_scene_desc, material_id, 40);
OFF(rt_scene_desc, n_materials, 48); OFF(rt_scene_desc, mat_type, 56);     OFF(rt_scene_desc, mat_param, 64);
OFF(rt_scene_desc, mat_tex, 72);     OFF(rt_scene_desc, n_textures, 80);   OFF(rt_scene_desc, tex_type, 88);
OFF(rt_scene_desc, tex_params, 96);  OFF(rt_scene_desc, tex_children, 104);
_Static_assert(sizeof(rt_scene_ext) == 120, "rt_scene_ext size: update clojure/native.clj");
OFF(rt_scene_ext, struct_bytes, 0);  OFF(rt_scene_ext, n_boundary, 4);     OFF(rt_scene_ext, prim_type, 8);
OFF(rt_scene_ext, prim_params, 16);  OFF(rt_scene_ext, prim_aux, 24);      OFF(rt_scene_ext, prim_xform, 32);
OFF(rt_scene_ext, n_xforms, 40);     OFF(rt_scene_ext, xform_ops, 48);     OFF(rt_scene_ext, xform_params, 56);
OFF(rt_scene_ext, tie_rule, 64);     OFF(rt_scene_ext, perlin_vectors, 72); OFF(rt_scene_ext, perlin_perm, 80);
OFF(rt_scene_ext, n_images, 88);     OFF(rt_scene_ext, image_wh, 96);      OFF(rt_scene_ext, image_offset, 104);
OFF(rt_scene_ext, image_rgb, 112);
_Static_assert(sizeof(rt_path_bounce) == 40, "rt_path_bounce size");
OFF(rt_path_bounce, o, 0); OFF(rt_path_bounce, time, 12); OFF(rt_path_bounce, d, 16); OFF(rt_path_bounce, hit_id, 28);
OFF(rt_path_bounce, t, 32);
_Static_assert(RT_CTR_COUNT == 24 && RT_ABI_VERSION == 2, "counter block / ABI version: update clojure/native.clj");

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != RT_OK) {                                                                      \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, rt_last_error(ctx));                   \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

int main(void) {
    rt_ctx* ctx = NULL;
    int dev = 0;
    int rc = rt_create(&ctx, &dev, 1);
    if (rc == RT_ERR_NODEVICE) {
        printf("nodevice: %s\n", rt_last_error(NULL));
        return 3;
    }
    if (rc != RT_OK) {
        fprintf(stderr, "rt_create -> %d: %s\n", rc, rt_last_error(NULL));
        return 1;
    }
    if (rt_abi_version() != RT_ABI_VERSION) return 1;

    /* world: sky dome (DiffuseLight, constant white), ground r=100 (Lambertian grey), ball r=0.5 (Lambertian red) */
    const float center0_r[12] = {0, 0, 0, 1000, 0, -100.5f, -1, 100, 0, 0, -1, 0.5f};
    const uint32_t flags[3] = {0, 0, 0};
    const int32_t material_id[3] = {0, 1, 2};
    const int32_t mat_type[3] = {RT_MAT_DIFFUSE_LIGHT, RT_MAT_LAMBERTIAN, RT_MAT_LAMBERTIAN};
    const float mat_param[3] = {0, 0, 0};
    const int32_t mat_tex[3] = {0, 1, 2};
    const int32_t tex_type[3] = {RT_TEX_CONSTANT, RT_TEX_CONSTANT, RT_TEX_CONSTANT};
    float tex_params[36];
    const int32_t tex_children[6] = {-1, -1, -1, -1, -1, -1};
    memset(tex_params, 0, sizeof tex_params);
    tex_params[0] = tex_params[1] = tex_params[2] = 1.0f;
    tex_params[12] = tex_params[13] = tex_params[14] = 0.5f;
    tex_params[24] = 0.7f; tex_params[25] = 0.3f; tex_params[26] = 0.3f;
    rt_scene_desc sc;
    memset(&sc, 0, sizeof sc);
    sc.n_spheres = 3; sc.center0_r = center0_r; sc.sphere_flags = flags; sc.material_id = material_id;
    sc.n_materials = 3; sc.mat_type = mat_type; sc.mat_param = mat_param; sc.mat_tex = mat_tex;
    sc.n_textures = 3; sc.tex_type = tex_type; sc.tex_params = tex_params; sc.tex_children = tex_children;
    CHECK(rt_set_scene(ctx, &sc));

    /* pinhole camera at the origin looking down -z (camera.clj:18-33 with vfov 90, aspect 2) */
    float cam[24];
    memset(cam, 0, sizeof cam);
    cam[3] = -2; cam[4] = -1; cam[5] = -1;   /* lleft */
    cam[6] = 4;                              /* horiz */
    cam[10] = 2;                             /* vert  */
    CHECK(rt_set_camera(ctx, RT_CAM_PINHOLE, cam));

    /* the ray from the origin down -z hits the ball at t = 0.5 exactly; straight up it ends on the sky dome at 1000 */
    const float o[6] = {0, 0, 0, 0, 0, 0}, d[6] = {0, 0, -1, 0, 1, 0};
    double t[2];
    int32_t id[2];
    CHECK(rt_trace_primary(ctx, 2, o, d, NULL, 0.001, 3.4028234663852886e38, t, id));
    if (id[0] != 2 || t[0] != 0.5 || id[1] != 0 || t[1] != 1000.0) {
        fprintf(stderr, "trace: id %d t %.17g, id %d t %.17g\n", id[0], t[0], id[1], t[1]);
        return 1;
    }

    enum { NX = 64, NY = 32, NS = 16 };
    float* lin = (float*)malloc(sizeof(float) * NX * NY * 3);
    uint8_t* rgb = (uint8_t*)malloc(NX * NY * 3);
    for (int variant = 0; variant < 2; ++variant) {
        CHECK(rt_render(ctx, NX, NY, NS, 50, 7u, variant, lin, rgb));
        /* top row of the 8-bit image = sky (white), centre = the red ball lit by the sky */
        const uint8_t* top = rgb + (NX / 2) * 3;
        const uint8_t* mid = rgb + ((NY / 2) * NX + NX / 2) * 3;
        if (top[0] < 250 || top[1] < 250 || top[2] < 250 || !(mid[0] > mid[1] + 20) || mid[0] < 60) {
            fprintf(stderr, "variant %d image: top %d %d %d centre %d %d %d\n", variant, top[0], top[1], top[2], mid[0], mid[1], mid[2]);
            return 1;
        }
    }
    /* the same frame into page-locked buffers (rt_host_alloc): written by the device directly, same 8-bit image
     * (up to a rounding edge: the float sums are accumulated in a different order from run to run) */
    float* plin = NULL;
    uint8_t* prgb = NULL;
    CHECK(rt_host_alloc(ctx, sizeof(float) * NX * NY * 3, (void**)&plin));
    CHECK(rt_host_alloc(NULL, NX * NY * 3, (void**)&prgb));
    memset(prgb, 0, NX * NY * 3);
    CHECK(rt_render(ctx, NX, NY, NS, 50, 7u, 1, plin, prgb));
    int differ = 0;
    for (int q = 0; q < NX * NY * 3; ++q) differ += abs((int)prgb[q] - (int)rgb[q]) > 1;
    if (differ) { fprintf(stderr, "page-locked output differs in %d bytes\n", differ); return 1; }
    CHECK(rt_host_free(ctx, plin));
    CHECK(rt_host_free(NULL, prgb));
    uint64_t ctr[RT_CTR_COUNT];
    CHECK(rt_get_counters(ctx, ctr));
    if (ctr[RT_CTR_SAMPLES] != 3u * NX * NY * NS || ctr[RT_CTR_SPHERE_TESTS] != ctr[RT_CTR_RAYS] * 3u) return 1;
    /* error behaviour: nonsense arguments come back as a status and a message, never a crash */
    if (rt_render(ctx, 0, NY, NS, 50, 1u, 1, lin, rgb) != RT_ERR_ARG || strlen(rt_last_error(ctx)) == 0) return 1;
    free(lin);
    free(rgb);
    rt_destroy(ctx);
    printf("ok: t = %.3f / %.1f, %llu rays\n", t[0], t[1], (unsigned long long)ctr[RT_CTR_RAYS]);
    return 0;
}
