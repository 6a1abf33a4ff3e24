//! Raw bindings to `include/bendy2d_b200.h` (ABI version 1).  UNVERIFIED: never compiled here.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_float, c_int, c_void};

#[repr(C)]
pub struct bendy_solver {
    _private: [u8; 0],
}

pub const BENDY_OK: c_int = 0;
pub const BENDY_ERR_ARG: c_int = -1;
pub const BENDY_ERR_LINK: c_int = -2;
pub const BENDY_ERR_CUDA: c_int = -3;
pub const BENDY_ERR_UNSUPPORTED: c_int = -4;
pub const BENDY_ERR_NO_DEVICE: c_int = -5;
pub const BENDY_LINKS_COLOURED: c_int = 0;
pub const BENDY_LINKS_REFERENCE_ORDER: c_int = 1;

#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct bendy_schedule_info {
    pub n_partitions: u32,
    pub n_local_colours: u32,
    pub n_global_colours: u32,
    pub n_local_links: u32,
    pub n_global_links: u32,
    pub n_poly_partitions: u32,
    pub kernels_per_substep: u32,
    pub n_priority_partitions: u32,
}

extern "C" {
    pub fn bendy_create(device: c_int) -> *mut bendy_solver;
    pub fn bendy_destroy(s: *mut bendy_solver);
    pub fn bendy_clone(s: *mut bendy_solver) -> *mut bendy_solver;
    pub fn bendy_last_error(s: *const bendy_solver) -> *const c_char;
    pub fn bendy_abi_version() -> c_int;
    pub fn bendy_save_snapshot(s: *mut bendy_solver, path: *const c_char) -> c_int;
    pub fn bendy_load_snapshot(path: *const c_char, device: c_int) -> *mut bendy_solver;
    pub fn bendy_get_last_update_args(s: *const bendy_solver, dt_g_bounds7: *mut c_float, valid: *mut c_int) -> c_int;

    pub fn bendy_add_particles(s: *mut bendy_solver, pos_xy: *const c_float, n: usize) -> c_int;
    pub fn bendy_add_circles(s: *mut bendy_solver, pos_xy: *const c_float, prev_xy: *const c_float,
                             acc_xy: *const c_float, radius: *const c_float, n: usize) -> c_int;
    pub fn bendy_add_polygon(s: *mut bendy_solver, pos_xy: *const c_float, prev_xy: *const c_float,
                             acc_xy: *const c_float, nv: usize, link_ab: *const u32, link_len: *const c_float,
                             nl: usize, is_static: c_int, cx: c_float, cy: c_float) -> c_int;
    pub fn bendy_add_particle_links(s: *mut bendy_solver, ab: *const u32, len: *const c_float, n: usize) -> c_int;
    pub fn bendy_add_circle_links(s: *mut bendy_solver, ab: *const u32, len: *const c_float, n: usize) -> c_int;

    pub fn bendy_update(s: *mut bendy_solver, dt: c_float, gx: c_float, gy: c_float, bx: c_float, by: c_float,
                        bw: c_float, bh: c_float) -> c_int;
    pub fn bendy_update_n(s: *mut bendy_solver, n: u32, dt: c_float, gx: c_float, gy: c_float, bx: c_float,
                          by: c_float, bw: c_float, bh: c_float) -> c_int;
    pub fn bendy_synchronize(s: *mut bendy_solver) -> c_int;

    pub fn bendy_particle_len(s: *const bendy_solver) -> usize;
    pub fn bendy_circle_len(s: *const bendy_solver) -> usize;
    pub fn bendy_polygon_len(s: *const bendy_solver) -> usize;
    pub fn bendy_particle_link_len(s: *const bendy_solver) -> usize;
    pub fn bendy_circle_link_len(s: *const bendy_solver) -> usize;
    pub fn bendy_polygon_point_len(s: *const bendy_solver, polygon: usize) -> usize;
    pub fn bendy_polygon_link_len(s: *const bendy_solver, polygon: usize) -> usize;
    pub fn bendy_read_particles(s: *mut bendy_solver, first: usize, n: usize, pos_xy: *mut c_float,
                                prev_xy: *mut c_float) -> c_int;
    pub fn bendy_read_circles(s: *mut bendy_solver, first: usize, n: usize, pos_xy: *mut c_float,
                              prev_xy: *mut c_float, radius: *mut c_float) -> c_int;
    pub fn bendy_read_polygon(s: *mut bendy_solver, polygon: usize, pos_xy: *mut c_float, prev_xy: *mut c_float,
                              center_xy: *mut c_float, is_static: *mut c_int) -> c_int;
    pub fn bendy_read_particle_links(s: *const bendy_solver, first: usize, n: usize, ab: *mut u32,
                                     len: *mut c_float) -> c_int;
    pub fn bendy_read_circle_links(s: *const bendy_solver, first: usize, n: usize, ab: *mut u32,
                                   len: *mut c_float) -> c_int;
    pub fn bendy_read_polygon_links(s: *const bendy_solver, polygon: usize, ab: *mut u32, len: *mut c_float) -> c_int;

    pub fn bendy_set_sub_steps(s: *mut bendy_solver, n: u16) -> c_int;
    pub fn bendy_write_particles(s: *mut bendy_solver, first: usize, n: usize, pos_xy: *const c_float,
                                 prev_xy: *const c_float) -> c_int;
    pub fn bendy_set_particle_radius(s: *mut bendy_solver, r: c_float) -> c_int;
    pub fn bendy_set_grid_cell(s: *mut bendy_solver, h: c_float) -> c_int;
    pub fn bendy_set_polygon_contact(s: *mut bendy_solver, on: c_int) -> c_int;
    pub fn bendy_set_particle_inv_mass(s: *mut bendy_solver, first: usize, n: usize, k: *const c_float) -> c_int;
    pub fn bendy_set_circle_inv_mass(s: *mut bendy_solver, first: usize, n: usize, k: *const c_float) -> c_int;
    pub fn bendy_set_plan_params(s: *mut bendy_solver, pack_points: u32, max_points: u32) -> c_int;
    pub fn bendy_set_link_schedule(s: *mut bendy_solver, mode: c_int) -> c_int;

    pub fn bendy_get_schedule_info(s: *mut bendy_solver, out: *mut bendy_schedule_info) -> c_int;
    pub fn bendy_get_link_order(s: *mut bendy_solver, perm: *mut u32, n: usize) -> c_int;
    pub fn bendy_get_point_rank(s: *mut bendy_solver, rank: *mut u32, n: usize) -> c_int;
    pub fn bendy_get_grid(s: *mut bendy_solver, bx: c_float, by: c_float, bw: c_float, bh: c_float,
                          ox: *mut c_float, oy: *mut c_float, inv_h: *mut c_float, nx: *mut c_int,
                          ny: *mut c_int) -> c_int;

    pub fn bendy_set_profiling(s: *mut bendy_solver, profile: c_int) -> c_int;
    pub fn bendy_get_kernel_times(s: *mut bendy_solver, ms: *mut f64, launches: *mut u64, n_classes: c_int,
                                  reset: c_int) -> c_int;
    pub fn bendy_launch_count(s: *const bendy_solver) -> u64;
    pub fn bendy_get_stats(s: *mut bendy_solver, out: *mut u64, n: c_int) -> c_int;
    pub fn bendy_timer_start(s: *mut bendy_solver) -> c_int;
    pub fn bendy_timer_stop(s: *mut bendy_solver, ms: *mut c_float) -> c_int;
    pub fn bendy_get_stream(s: *const bendy_solver) -> *mut c_void;
    pub fn bendy_get_device(s: *const bendy_solver) -> c_int;
    pub fn bendy_get_device_buffers(s: *mut bendy_solver, pos: *mut *mut c_void, prev: *mut *mut c_void,
                                    n_points: *mut usize) -> c_int;

    pub fn bendy_halo_configure(s: *mut bendy_solver, ghost_cap: u32, x_left: c_float, x_right: c_float,
                                stray_left: c_float, stray_right: c_float) -> c_int;
    pub fn bendy_set_grid_window(s: *mut bendy_solver, x0: c_float, x1: c_float) -> c_int;
    pub fn bendy_strip_set_cross_links(s: *mut bendy_solver, n: usize, mine: *const u32, slot: *const u32, i_am_a: *const u8,
                                       len: *const c_float, n_colours: u32, colour_start: *const u32, n_send_left: usize,
                                       send_left: *const u32, n_send_right: usize, send_right: *const u32,
                                       n_recv_left: usize, n_recv_right: usize) -> c_int;
    pub fn bendy_nccl_unique_id(out128: *mut c_void) -> c_int;
    pub fn bendy_halo_comm_nccl(s: *mut bendy_solver, unique_id128: *const c_void, rank: c_int, world: c_int) -> c_int;
    pub fn bendy_halo_connect_local(left: *mut bendy_solver, right: *mut bendy_solver) -> c_int;
    pub fn bendy_update_group(group: *mut *mut bendy_solver, n: c_int, n_updates: u32, dt: c_float, gx: c_float,
                              gy: c_float, bx: c_float, by: c_float, bw: c_float, bh: c_float) -> c_int;
    pub fn bendy_halo_stats(s: *mut bendy_solver, sent_left: *mut u32, sent_right: *mut u32, overflow: *mut u32,
                            strayed: *mut u32) -> c_int;

    pub fn bendy_plan_links(n_points: usize, ab: *const u32, n_links: usize, pack_points: u32, max_points: u32,
                            rank: *mut u32, perm: *mut u32, link_colour: *mut u32, link_partition: *mut u32,
                            info: *mut bendy_schedule_info) -> c_int;
    pub fn bendy_plan_links_scheduled(n_points: usize, ab: *const u32, n_links: usize, pack_points: u32, max_points: u32,
                                      link_schedule: c_int, rank: *mut u32, perm: *mut u32, link_colour: *mut u32,
                                      link_partition: *mut u32, info: *mut bendy_schedule_info) -> c_int;
    // test hook: normalize() of link.rs:24 / circle.rs:37 as the kernels compute it, for n host triples
    pub fn bendy_debug_normalize(device: c_int, dx: *const f32, dy: *const f32, norm: *const f32, n: usize,
                                 nx: *mut f32, ny: *mut f32) -> c_int;
}
