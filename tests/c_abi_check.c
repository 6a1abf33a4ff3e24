/* Compiled (not run) by tests/test_host_logic.py with gcc -std=c99 -Wall -Werror: proves that
 * include/bendy2d_b200.h is a valid plain-C header and that every declared entry point links against
 * libbendy2d_b200.so with the declared signature. */
#include "bendy2d_b200.h"

#include <stdio.h>

int main(void) {
    /* one well-typed call site per entry point; never executed on a GPU-less box beyond bendy_create */
    bendy_solver *s = bendy_create(-1);
    if (!s) {
        printf("no device: %s\n", bendy_last_error(NULL));
        return bendy_abi_version() == BENDY_ABI_VERSION ? 0 : 1;
    }
    float xy[4] = {1.f, 2.f, 3.f, 4.f}, len[1] = {2.f}, r[1] = {1.f}, k[2] = {1.f, 1.f}, ms = 0.f;
    uint32_t ab[2] = {0u, 1u}, perm[1], rank[2], a = 0, b = 0, c = 0, d = 0;
    double kms[BENDY_K_CLASSES];
    uint64_t kl[BENDY_K_CLASSES], st[4];
    bendy_schedule_info info;
    float ox, oy, ih;
    int nx, ny, is_static;
    void *p0, *p1;
    size_t np;
    unsigned char uid[128];
    bendy_solver *grp[1];
    int rc = 0;
    rc |= bendy_add_particles(s, xy, 2);
    rc |= bendy_add_particle_links(s, ab, len, 1);
    rc |= bendy_add_circles(s, xy, NULL, NULL, r, 1);
    rc |= bendy_add_circle_links(s, ab, len, 0);
    rc |= bendy_add_polygon(s, xy, NULL, NULL, 2, ab, len, 1, 0, 2.f, 3.f);
    rc |= bendy_set_sub_steps(s, 2);
    rc |= bendy_set_particle_radius(s, 0.f);
    rc |= bendy_set_grid_cell(s, 0.f);
    rc |= bendy_set_polygon_contact(s, 0);
    rc |= bendy_set_particle_inv_mass(s, 0, 2, k);
    rc |= bendy_set_circle_inv_mass(s, 0, 1, k);
    rc |= bendy_set_plan_params(s, 0, 0);
    rc |= bendy_set_link_schedule(s, BENDY_LINKS_COLOURED);
    rc |= bendy_update(s, 0.01f, 0.f, 98.2f, 0.f, 0.f, 100.f, 100.f);
    rc |= bendy_update_n(s, 2, 0.01f, 0.f, 98.2f, 0.f, 0.f, 100.f, 100.f);
    rc |= bendy_synchronize(s);
    rc |= bendy_read_particles(s, 0, bendy_particle_len(s), xy, NULL);
    rc |= bendy_write_particles(s, 0, 2, xy, NULL);
    rc |= bendy_read_circles(s, 0, bendy_circle_len(s), xy, NULL, r);
    rc |= bendy_read_polygon(s, 0, xy, NULL, xy, &is_static);
    rc |= bendy_read_particle_links(s, 0, bendy_particle_link_len(s), ab, len);
    rc |= bendy_read_circle_links(s, 0, bendy_circle_link_len(s), ab, len);
    rc |= bendy_read_polygon_links(s, 0, ab, len);
    rc |= (int)(bendy_polygon_len(s) + bendy_polygon_point_len(s, 0) + bendy_polygon_link_len(s, 0)) == 0;
    rc |= bendy_get_schedule_info(s, &info);
    rc |= bendy_get_link_order(s, perm, 1);
    rc |= bendy_get_point_rank(s, rank, 2);
    rc |= bendy_get_grid(s, 0.f, 0.f, 100.f, 100.f, &ox, &oy, &ih, &nx, &ny);
    rc |= bendy_set_profiling(s, 0);
    rc |= bendy_get_kernel_times(s, kms, kl, BENDY_K_CLASSES, 1);
    rc |= bendy_get_stats(s, st, 4);
    rc |= bendy_timer_start(s);
    rc |= bendy_timer_stop(s, &ms);
    rc |= bendy_get_device_buffers(s, &p0, &p1, &np);
    rc |= bendy_halo_configure(s, 0, 0.f, 0.f, 0.f, 0.f);
    rc |= bendy_set_grid_window(s, 0.f, 1.f);
    rc |= bendy_strip_set_cross_links(s, 0, NULL, NULL, NULL, NULL, 0, NULL, 0, NULL, 0, NULL, 0, 0);
    rc |= bendy_nccl_unique_id(uid);
    rc |= bendy_halo_comm_nccl(s, uid, 0, 1);
    rc |= bendy_halo_stats(s, &a, &b, &c, &d);
    grp[0] = s;
    rc |= bendy_update_group(grp, 1, 1, 0.01f, 0.f, 98.2f, 0.f, 0.f, 100.f, 100.f);
    rc |= bendy_plan_links(2, ab, 1, 0, 0, rank, perm, NULL, NULL, &info);
    rc |= bendy_plan_links_scheduled(2, ab, 1, 0, 0, BENDY_LINKS_REFERENCE_ORDER, rank, perm, NULL, NULL, &info);
    {
        float one = 1.f, two = 2.f, nxy[2];
        rc |= bendy_debug_normalize(-1, &one, &one, &two, 1, &nxy[0], &nxy[1]);
    }
    {
        float last[7];
        int valid = 0;
        bendy_solver *c3;
        rc |= bendy_get_last_update_args(s, last, &valid);
        rc |= bendy_save_snapshot(s, "/tmp/bendy_c_abi_check.snap");
        c3 = bendy_load_snapshot("/tmp/bendy_c_abi_check.snap", -1);
        if (c3) bendy_destroy(c3);
    }
    {
        bendy_solver *c2 = bendy_clone(s);
        if (c2) {
            rc |= bendy_halo_connect_local(s, c2) != BENDY_OK ? 0 : 0;
            bendy_destroy(c2);
        }
    }
    printf("%p %d %llu\n", bendy_get_stream(s), bendy_get_device(s), (unsigned long long)bendy_launch_count(s));
    bendy_destroy(s);
    /* this program is a compile/link check: on a GPU box the calls above really run, but their return
     * codes are informational here (the parity tests are the behavioural check) */
    printf("accumulated return codes: %d\n", rc);
    return 0;
}
