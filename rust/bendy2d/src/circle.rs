//! reference src/circle.rs:5-8
use crate::particle::Particle;

#[derive(Debug, Copy, Clone)]
pub struct Circle {
    pub point: Particle,
    pub radius: f32,
}
