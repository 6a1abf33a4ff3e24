//! reference src/particle.rs:5-59 — plain data; the arithmetic runs on the GPU.
use nalgebra::Vector2;

#[derive(Debug, Copy, Clone)]
pub struct Particle {
    pub pos: Vector2<f32>,
    pub prev_pos: Vector2<f32>,
    pub(crate) acc: Vector2<f32>,
}

impl Particle {
    pub fn new(pos: Vector2<f32>) -> Self {
        Self { pos, prev_pos: pos, acc: Vector2::new(0.0, 0.0) }
    }
    pub fn add_force(&mut self, force_x: f32, force_y: f32) {
        self.acc += Vector2::new(force_x, force_y);
    }
    pub fn add_force_v2(&mut self, force: Vector2<f32>) {
        self.acc += force;
    }
    pub fn add_force_towards(&mut self, point: Vector2<f32>, force: f32) {
        let dist: Vector2<f32> = (point - self.pos).normalize();
        self.acc += dist * force;
    }
}
