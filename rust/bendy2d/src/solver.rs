//! Drop-in for the reference `Solver` (reference src/solver.rs:13-116) over libbendy2d_b200.
//! UNVERIFIED: written against include/bendy2d_b200.h, never compiled (no Rust toolchain in the image).
//!
//! Differences a caller can observe:
//!  * `get_*` take `&mut self`-free borrows in the reference; here the AoS mirrors are refreshed lazily
//!    behind `RefCell`-free interior state, so the getters take `&mut self` (a one-line change at call
//!    sites) — or call `sync()` once and use the `&self` getters;
//!  * an invalid link panics at `add_*_link` (the reference panics inside `update`, link.rs:19-21);
//!  * `set_sub_steps`, `set_particle_radius`, `set_polygon_contact` are additive.
use crate::circle::Circle;
use crate::link::{CircleLink, Link, ParticleLink};
use crate::particle::Particle;
use crate::polygon::Polygon;
use bendy2d_sys as sys;
use nalgebra::Vector2;
use std::ffi::CStr;

#[derive(Debug, Copy, Clone)]
pub struct Bounds {
    pub pos: Vector2<f32>,
    pub size: Vector2<f32>,
}

pub struct Solver {
    pub gravity: Vector2<f32>,
    pub bounds: Bounds,
    pub bounds_active: bool, // never read by the reference either (solver.rs:155-165)
    handle: *mut sys::bendy_solver,
    particles: Vec<Particle>,
    particle_links: Vec<ParticleLink>,
    circles: Vec<Circle>,
    circle_links: Vec<CircleLink>,
    polygons: Vec<Polygon>,
    stale: bool, // device state is newer than the host mirrors
}

// a handle owns a CUDA stream: movable between threads, not shareable
unsafe impl Send for Solver {}

impl Solver {
    pub fn new() -> Self {
        let handle = unsafe { sys::bendy_create(-1) };
        assert!(!handle.is_null(), "bendy_create failed: {}", Self::err(std::ptr::null()));
        Self {
            gravity: Vector2::new(0.0, 98.2),
            bounds: Bounds { pos: Vector2::new(0.0, 0.0), size: Vector2::new(100.0, 100.0) },
            bounds_active: true,
            handle,
            particles: Vec::new(),
            particle_links: Vec::new(),
            circles: Vec::new(),
            circle_links: Vec::new(),
            polygons: Vec::new(),
            stale: false,
        }
    }

    fn err(h: *const sys::bendy_solver) -> String {
        unsafe { CStr::from_ptr(sys::bendy_last_error(h)).to_string_lossy().into_owned() }
    }
    fn ck(&self, rc: i32) {
        if rc != sys::BENDY_OK {
            panic!("bendy2d_b200: {}", Self::err(self.handle));
        }
    }

    pub fn add_particle(&mut self, pos: Vector2<f32>) {
        let xy = [pos.x, pos.y];
        self.ck(unsafe { sys::bendy_add_particles(self.handle, xy.as_ptr(), 1) });
        self.particles.push(Particle::new(pos));
    }
    pub fn add_circle(&mut self, circle: Circle) {
        let p = [circle.point.pos.x, circle.point.pos.y];
        let q = [circle.point.prev_pos.x, circle.point.prev_pos.y];
        let a = [circle.point.acc.x, circle.point.acc.y];
        self.ck(unsafe { sys::bendy_add_circles(self.handle, p.as_ptr(), q.as_ptr(), a.as_ptr(), &circle.radius, 1) });
        self.circles.push(circle);
    }
    pub fn add_polygon(&mut self, polygon: Polygon) {
        let flat = |f: &dyn Fn(&Particle) -> Vector2<f32>| -> Vec<f32> {
            polygon.particles.iter().flat_map(|p| { let v = f(p); [v.x, v.y] }).collect()
        };
        let (pos, prev, acc) = (flat(&|p| p.pos), flat(&|p| p.prev_pos), flat(&|p| p.acc));
        let ab: Vec<u32> = polygon.particle_links.iter()
            .flat_map(|l| [l.link.particle_a as u32, l.link.particle_b as u32]).collect();
        let len: Vec<f32> = polygon.particle_links.iter().map(|l| l.link.target_distance).collect();
        self.ck(unsafe {
            sys::bendy_add_polygon(self.handle, pos.as_ptr(), prev.as_ptr(), acc.as_ptr(), polygon.particles.len(),
                                   ab.as_ptr(), len.as_ptr(), len.len(), polygon.is_static as i32,
                                   polygon.center.x, polygon.center.y)
        });
        self.polygons.push(polygon);
    }
    pub fn add_particle_link(&mut self, link: ParticleLink) {
        let ab = [link.link.particle_a as u32, link.link.particle_b as u32];
        self.ck(unsafe { sys::bendy_add_particle_links(self.handle, ab.as_ptr(), &link.link.target_distance, 1) });
        self.particle_links.push(link);
    }
    pub fn add_circle_link(&mut self, link: CircleLink) {
        let ab = [link.link.particle_a as u32, link.link.particle_b as u32];
        self.ck(unsafe { sys::bendy_add_circle_links(self.handle, ab.as_ptr(), &link.link.target_distance, 1) });
        self.circle_links.push(link);
    }

    pub fn get_particle_len(&self) -> usize { self.particles.len() }
    pub fn get_circles_len(&self) -> usize { self.circles.len() }
    pub fn get_polygons_len(&self) -> usize { self.polygons.len() }
    pub fn get_particle_links(&self) -> &Vec<ParticleLink> { &self.particle_links }
    pub fn get_circle_links(&self) -> &Vec<CircleLink> { &self.circle_links }

    /// Refreshes the host AoS mirrors from the device (one D2H per class); no-op when current.
    pub fn sync(&mut self) {
        if !self.stale {
            return;
        }
        let n = self.particles.len();
        let (mut pos, mut prev) = (vec![0f32; 2 * n], vec![0f32; 2 * n]);
        self.ck(unsafe { sys::bendy_read_particles(self.handle, 0, n, pos.as_mut_ptr(), prev.as_mut_ptr()) });
        for (i, p) in self.particles.iter_mut().enumerate() {
            p.pos = Vector2::new(pos[2 * i], pos[2 * i + 1]);
            p.prev_pos = Vector2::new(prev[2 * i], prev[2 * i + 1]);
            p.acc = Vector2::new(0.0, 0.0);
        }
        let nc = self.circles.len();
        let (mut cp, mut cq, mut cr) = (vec![0f32; 2 * nc], vec![0f32; 2 * nc], vec![0f32; nc]);
        self.ck(unsafe { sys::bendy_read_circles(self.handle, 0, nc, cp.as_mut_ptr(), cq.as_mut_ptr(), cr.as_mut_ptr()) });
        for (i, c) in self.circles.iter_mut().enumerate() {
            c.point.pos = Vector2::new(cp[2 * i], cp[2 * i + 1]);
            c.point.prev_pos = Vector2::new(cq[2 * i], cq[2 * i + 1]);
        }
        for (k, poly) in self.polygons.iter_mut().enumerate() {
            let nv = poly.particles.len();
            let (mut pp, mut pq, mut cen, mut st) = (vec![0f32; 2 * nv], vec![0f32; 2 * nv], [0f32; 2], 0i32);
            let rc = unsafe { sys::bendy_read_polygon(self.handle, k, pp.as_mut_ptr(), pq.as_mut_ptr(), cen.as_mut_ptr(), &mut st) };
            assert_eq!(rc, sys::BENDY_OK);
            for (i, p) in poly.particles.iter_mut().enumerate() {
                p.pos = Vector2::new(pp[2 * i], pp[2 * i + 1]);
                p.prev_pos = Vector2::new(pq[2 * i], pq[2 * i + 1]);
            }
            poly.center = Vector2::new(cen[0], cen[1]);
        }
        self.stale = false;
    }

    pub fn get_particles(&mut self) -> &Vec<Particle> { self.sync(); &self.particles }
    pub fn get_circles(&mut self) -> &Vec<Circle> { self.sync(); &self.circles }
    pub fn get_polygons(&mut self) -> &Vec<Polygon> { self.sync(); &self.polygons }
    pub fn get_particle(&mut self, index: usize) -> Option<&Particle> { self.sync(); self.particles.get(index) }
    pub fn get_circle(&mut self, index: usize) -> Option<&Circle> { self.sync(); self.circles.get(index) }
    pub fn get_polygon(&mut self, index: usize) -> Option<&Polygon> { self.sync(); self.polygons.get(index) }

    /// reference solver.rs:106-116; asynchronous: the next getter synchronises.
    pub fn update(&mut self, dt: f32) {
        self.ck(unsafe {
            sys::bendy_update(self.handle, dt, self.gravity.x, self.gravity.y, self.bounds.pos.x, self.bounds.pos.y,
                              self.bounds.size.x, self.bounds.size.y)
        });
        self.stale = true;
    }

    // ---- additive
    pub fn set_sub_steps(&mut self, n: u16) { self.ck(unsafe { sys::bendy_set_sub_steps(self.handle, n) }); }
    pub fn set_particle_radius(&mut self, r: f32) { self.ck(unsafe { sys::bendy_set_particle_radius(self.handle, r) }); }
    pub fn set_polygon_contact(&mut self, on: bool) { self.ck(unsafe { sys::bendy_set_polygon_contact(self.handle, on as i32) }); }

    /// The state `clone()` copies, as one flat file (format: bendy2d_b200/snapshot.py).  Loading goes through
    /// `bendy2d_sys::bendy_load_snapshot`; rebuilding this façade's AoS mirrors from a loaded handle is not
    /// written yet.
    pub fn save_snapshot(&mut self, path: &std::path::Path) {
        let c = std::ffi::CString::new(path.to_string_lossy().as_bytes()).expect("path contains a NUL byte");
        self.ck(unsafe { sys::bendy_save_snapshot(self.handle, c.as_ptr()) });
    }
}

impl Clone for Solver {
    fn clone(&self) -> Self {
        let handle = unsafe { sys::bendy_clone(self.handle) };
        assert!(!handle.is_null(), "bendy_clone failed: {}", Self::err(self.handle));
        Self {
            gravity: self.gravity, bounds: self.bounds, bounds_active: self.bounds_active, handle,
            particles: self.particles.clone(), particle_links: self.particle_links.clone(),
            circles: self.circles.clone(), circle_links: self.circle_links.clone(), polygons: self.polygons.clone(),
            stale: true,
        }
    }
}

impl Drop for Solver {
    fn drop(&mut self) {
        unsafe { sys::bendy_destroy(self.handle) }
    }
}

// keep `Link` in the public path like the reference's `use crate::link::...`
#[allow(dead_code)]
fn _link_is_used(_: Link) {}
