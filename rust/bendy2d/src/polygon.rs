//! reference src/polygon.rs:8-123 — data + the two host-side constructors (same link tables).
use crate::link::{Link, ParticleLink};
use crate::particle::Particle;
use nalgebra::Vector2;

#[derive(Debug, Clone)]
pub struct Polygon {
    pub particles: Vec<Particle>,
    pub particle_links: Vec<ParticleLink>,
    pub is_static: bool,
    pub center: Vector2<f32>,
    pub scale: f32,
}

fn link_between(particles: &[Particle], mut a_id: usize, mut b_id: usize) -> ParticleLink {
    if a_id > b_id {
        std::mem::swap(&mut a_id, &mut b_id);
    }
    let dist_vec = particles[a_id].pos - particles[b_id].pos;
    ParticleLink { link: Link { particle_a: a_id, particle_b: b_id, target_distance: dist_vec.magnitude() } }
}

impl Polygon {
    /// reference polygon.rs:17-82
    pub fn circle(radius: f32, pos: Vector2<f32>, point_count: usize, is_static: bool) -> Self {
        let mut particles = Vec::new();
        let mut center = Vector2::new(0.0, 0.0);
        let mut angle = 0.0f32;
        for _ in 0..point_count {
            let point = Particle::new(pos + Vector2::new(radius * f32::cos(angle), radius * f32::sin(angle)));
            particles.push(point);
            center += point.pos;
            angle += 2.0 * std::f32::consts::PI / point_count as f32;
        }
        center /= point_count as f32;
        let mut particle_links = Vec::new();
        for i in 0..point_count {
            particle_links.push(link_between(&particles, i, (i + 2 * point_count / 3) % point_count));
            particle_links.push(link_between(&particles, i, (i + point_count / 3) % point_count));
        }
        Self { particles, particle_links, is_static, center, scale: 1.0 }
    }

    /// reference polygon.rs:84-123
    pub fn new(points: Vec<Vector2<f32>>, is_static: bool) -> Self {
        let mut particles = Vec::new();
        let mut center = Vector2::new(0.0, 0.0);
        for point in &points {
            let particle = Particle::new(*point);
            particles.push(particle);
            center += particle.pos;
        }
        center /= points.len() as f32;
        let n = particles.len();
        let particle_links = (0..n).map(|i| link_between(&particles, i, (i + 1) % n)).collect();
        Self { particles, particle_links, is_static, center, scale: 1.0 }
    }
}
