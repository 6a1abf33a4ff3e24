//! reference src/link.rs:5-15,30-33
#[derive(Debug, Copy, Clone)]
pub struct Link {
    // particle_a needs to be lower than particle_b
    pub particle_a: usize,
    pub particle_b: usize,
    pub target_distance: f32,
}
#[derive(Debug, Copy, Clone)]
pub struct ParticleLink {
    pub link: Link,
}
#[derive(Debug, Copy, Clone)]
pub struct CircleLink {
    pub link: Link,
}
