//! Drop-in for the reference `Solver` (reference src/solver.rs:13-116) over libbendy2d_b200.
//! UNVERIFIED: written against include/bendy2d_b200.h, never compiled (no Rust toolchain in the image).
//!
//! The getters keep the reference's signatures (`&self`, solver.rs:69-104): the AoS mirrors live in an
//! `UnsafeCell` and are refreshed lazily by the first getter after an `update`.  That is sound because
//!  * the mirrors only go stale in `update(&mut self)` / `add_*(&mut self)`, i.e. when no `&` borrow handed out by
//!    a getter can be alive, and a refresh happens at most once per staleness (later getters find them current
//!    and do not write);
//!  * `Solver` holds a raw handle, hence is `!Sync`: no second thread can call a getter concurrently.
//! `Debug` (solver.rs:19) is implemented by hand over the synchronised mirrors.
//!
//! Differences a caller can observe:
//!  * an invalid particle / circle link panics inside `update`, like the reference (link.rs:19-21); only a polygon's
//!    own links are checked when the polygon is added;
//!  * the host-side helpers the reference's own `update` is made of (`Particle::update`, `solve_bounds`,
//!    `Circle::solve_circle`, `ParticleLink::solve`, `Polygon::solve_*` ...) are not re-exported: that arithmetic
//!    runs on the device, and a host call on a mirror would be overwritten by the next refresh;
//!  * `set_sub_steps`, `set_particle_radius`, `set_polygon_contact`, `set_link_schedule`, `add_particles` (batch)
//!    are additive.
use crate::circle::Circle;
use crate::link::{CircleLink, Link, ParticleLink};
use crate::particle::Particle;
use crate::polygon::Polygon;
use bendy2d_sys as sys;
use nalgebra::Vector2;
use std::cell::UnsafeCell;
use std::ffi::CStr;

/// reference solver.rs:7-11 (never used there either)
#[derive(Debug, Copy, Clone, PartialEq, Eq)]
pub enum ColliderType {
    Particle,
    Circle,
    Polygon,
}

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
    particle_links: Vec<ParticleLink>,
    circle_links: Vec<CircleLink>,
    mirror: UnsafeCell<Mirror>,
}

/// host AoS copies of the device state, in the reference's types
#[derive(Clone)]
struct Mirror {
    particles: Vec<Particle>,
    circles: Vec<Circle>,
    polygons: Vec<Polygon>,
    stale: bool, // device state is newer than these copies
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
            particle_links: Vec::new(),
            circle_links: Vec::new(),
            mirror: UnsafeCell::new(Mirror { particles: Vec::new(), circles: Vec::new(), polygons: Vec::new(), stale: false }),
        }
    }
    fn m(&mut self) -> &mut Mirror {
        self.mirror.get_mut()
    }
    fn idx(i: usize) -> u32 {
        u32::try_from(i).expect("bendy2d_b200: link index does not fit 32 bits")
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
        self.m().particles.push(Particle::new(pos));
    }
    /// additive: many particles in one call across the FFI
    pub fn add_particles(&mut self, pos: &[Vector2<f32>]) {
        let xy: Vec<f32> = pos.iter().flat_map(|p| [p.x, p.y]).collect();
        self.ck(unsafe { sys::bendy_add_particles(self.handle, xy.as_ptr(), pos.len()) });
        self.m().particles.extend(pos.iter().map(|p| Particle::new(*p)));
    }
    pub fn add_circle(&mut self, circle: Circle) {
        let p = [circle.point.pos.x, circle.point.pos.y];
        let q = [circle.point.prev_pos.x, circle.point.prev_pos.y];
        let a = [circle.point.acc.x, circle.point.acc.y];
        self.ck(unsafe { sys::bendy_add_circles(self.handle, p.as_ptr(), q.as_ptr(), a.as_ptr(), &circle.radius, 1) });
        self.m().circles.push(circle);
    }
    pub fn add_polygon(&mut self, polygon: Polygon) {
        let flat = |f: &dyn Fn(&Particle) -> Vector2<f32>| -> Vec<f32> {
            polygon.particles.iter().flat_map(|p| { let v = f(p); [v.x, v.y] }).collect()
        };
        let (pos, prev, acc) = (flat(&|p| p.pos), flat(&|p| p.prev_pos), flat(&|p| p.acc));
        let ab: Vec<u32> = polygon.particle_links.iter()
            .flat_map(|l| [Self::idx(l.link.particle_a), Self::idx(l.link.particle_b)]).collect();
        let len: Vec<f32> = polygon.particle_links.iter().map(|l| l.link.target_distance).collect();
        self.ck(unsafe {
            sys::bendy_add_polygon(self.handle, pos.as_ptr(), prev.as_ptr(), acc.as_ptr(), polygon.particles.len(),
                                   ab.as_ptr(), len.as_ptr(), len.len(), polygon.is_static as i32,
                                   polygon.center.x, polygon.center.y)
        });
        self.m().polygons.push(polygon);
    }
    pub fn add_particle_link(&mut self, link: ParticleLink) {
        let ab = [Self::idx(link.link.particle_a), Self::idx(link.link.particle_b)];
        self.ck(unsafe { sys::bendy_add_particle_links(self.handle, ab.as_ptr(), &link.link.target_distance, 1) });
        self.particle_links.push(link);
    }
    pub fn add_circle_link(&mut self, link: CircleLink) {
        let ab = [Self::idx(link.link.particle_a), Self::idx(link.link.particle_b)];
        self.ck(unsafe { sys::bendy_add_circle_links(self.handle, ab.as_ptr(), &link.link.target_distance, 1) });
        self.circle_links.push(link);
    }

    pub fn get_particle_len(&self) -> usize { self.view().particles.len() }
    pub fn get_circles_len(&self) -> usize { self.view().circles.len() }
    pub fn get_polygons_len(&self) -> usize { self.view().polygons.len() }
    pub fn get_particle_links(&self) -> &Vec<ParticleLink> { &self.particle_links }
    pub fn get_circle_links(&self) -> &Vec<CircleLink> { &self.circle_links }

    /// The mirrors without a refresh (lengths and topology never go stale).
    fn view(&self) -> &Mirror {
        // SAFETY: see the module comment - no `&mut` to the mirror exists while `&self` is alive except inside
        // `synced()` below, which only writes when `stale` is set, i.e. before any getter has handed out a borrow.
        unsafe { &*self.mirror.get() }
    }

    /// The mirrors, refreshed from the device if an `update` ran since the last refresh (one D2H per class).
    fn synced(&self) -> &Mirror {
        // SAFETY: as above; `Solver: !Sync`, so this is the only thread in here.
        let m = unsafe { &mut *self.mirror.get() };
        if m.stale {
            let n = m.particles.len();
            let (mut pos, mut prev) = (vec![0f32; 2 * n], vec![0f32; 2 * n]);
            self.ck(unsafe { sys::bendy_read_particles(self.handle, 0, n, pos.as_mut_ptr(), prev.as_mut_ptr()) });
            for (i, p) in m.particles.iter_mut().enumerate() {
                p.pos = Vector2::new(pos[2 * i], pos[2 * i + 1]);
                p.prev_pos = Vector2::new(prev[2 * i], prev[2 * i + 1]);
                p.acc = Vector2::new(0.0, 0.0);
            }
            let nc = m.circles.len();
            let (mut cp, mut cq, mut cr) = (vec![0f32; 2 * nc], vec![0f32; 2 * nc], vec![0f32; nc]);
            self.ck(unsafe { sys::bendy_read_circles(self.handle, 0, nc, cp.as_mut_ptr(), cq.as_mut_ptr(), cr.as_mut_ptr()) });
            for (i, c) in m.circles.iter_mut().enumerate() {
                c.point.pos = Vector2::new(cp[2 * i], cp[2 * i + 1]);
                c.point.prev_pos = Vector2::new(cq[2 * i], cq[2 * i + 1]);
                c.point.acc = Vector2::new(0.0, 0.0);
            }
            for (k, poly) in m.polygons.iter_mut().enumerate() {
                let nv = poly.particles.len();
                let (mut pp, mut pq, mut cen, mut st) = (vec![0f32; 2 * nv], vec![0f32; 2 * nv], [0f32; 2], 0i32);
                self.ck(unsafe { sys::bendy_read_polygon(self.handle, k, pp.as_mut_ptr(), pq.as_mut_ptr(), cen.as_mut_ptr(), &mut st) });
                for (i, p) in poly.particles.iter_mut().enumerate() {
                    p.pos = Vector2::new(pp[2 * i], pp[2 * i + 1]);
                    p.prev_pos = Vector2::new(pq[2 * i], pq[2 * i + 1]);
                }
                poly.center = Vector2::new(cen[0], cen[1]);
            }
            m.stale = false;
        }
        m
    }

    // reference solver.rs:86-104: `&self`, borrows into solver-owned vectors
    pub fn get_particles(&self) -> &Vec<Particle> { &self.synced().particles }
    pub fn get_circles(&self) -> &Vec<Circle> { &self.synced().circles }
    pub fn get_polygons(&self) -> &Vec<Polygon> { &self.synced().polygons }
    pub fn get_particle(&self, index: usize) -> Option<&Particle> { self.synced().particles.get(index) }
    pub fn get_circle(&self, index: usize) -> Option<&Circle> { self.synced().circles.get(index) }
    pub fn get_polygon(&self, index: usize) -> Option<&Polygon> { self.synced().polygons.get(index) }

    /// reference solver.rs:106-116; asynchronous: the next getter synchronises.
    pub fn update(&mut self, dt: f32) {
        self.ck(unsafe {
            sys::bendy_update(self.handle, dt, self.gravity.x, self.gravity.y, self.bounds.pos.x, self.bounds.pos.y,
                              self.bounds.size.x, self.bounds.size.y)
        });
        self.m().stale = true;
    }

    // ---- additive
    pub fn set_sub_steps(&mut self, n: u16) { self.ck(unsafe { sys::bendy_set_sub_steps(self.handle, n) }); }
    pub fn set_particle_radius(&mut self, r: f32) { self.ck(unsafe { sys::bendy_set_particle_radius(self.handle, r) }); }
    pub fn set_polygon_contact(&mut self, on: bool) { self.ck(unsafe { sys::bendy_set_polygon_contact(self.handle, on as i32) }); }
    /// `true`: dependency-level colours, results equal the reference's insertion-order walk (solver.rs:143-146) bit
    /// for bit; `false` (default): greedy colouring, results equal the reference fed the links in the exported order.
    pub fn set_link_schedule(&mut self, reference_order: bool) {
        let mode = if reference_order { sys::BENDY_LINKS_REFERENCE_ORDER } else { sys::BENDY_LINKS_COLOURED };
        self.ck(unsafe { sys::bendy_set_link_schedule(self.handle, mode) });
    }

    /// The state `clone()` copies, as one flat file (format: bendy2d_b200/snapshot.py).  Loading goes through
    /// `bendy2d_sys::bendy_load_snapshot`; rebuilding this façade's AoS mirrors from a loaded handle is not
    /// written yet.
    pub fn save_snapshot(&self, path: &std::path::Path) {
        let c = std::ffi::CString::new(path.to_string_lossy().as_bytes()).expect("path contains a NUL byte");
        self.ck(unsafe { sys::bendy_save_snapshot(self.handle, c.as_ptr()) });
    }
}

impl Clone for Solver {
    fn clone(&self) -> Self {
        let handle = unsafe { sys::bendy_clone(self.handle) };
        assert!(!handle.is_null(), "bendy_clone failed: {}", Self::err(self.handle));
        let mut mirror = self.view().clone();
        mirror.stale = true; // the clone reads its own device state on the first getter
        Self {
            gravity: self.gravity, bounds: self.bounds, bounds_active: self.bounds_active, handle,
            particle_links: self.particle_links.clone(), circle_links: self.circle_links.clone(),
            mirror: UnsafeCell::new(mirror),
        }
    }
}

impl std::fmt::Debug for Solver {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        let m = self.synced();
        f.debug_struct("Solver")
            .field("gravity", &self.gravity)
            .field("bounds", &self.bounds)
            .field("bounds_active", &self.bounds_active)
            .field("particles", &m.particles)
            .field("circles", &m.circles)
            .field("polygons", &m.polygons)
            .field("particle_links", &self.particle_links)
            .field("circle_links", &self.circle_links)
            .finish()
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
