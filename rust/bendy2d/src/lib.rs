//! Drop-in façade: the reference's module layout (reference src/lib.rs:1-6).  UNVERIFIED sources.
pub mod circle;
pub mod link;
pub mod particle;
pub mod polygon;
pub mod solver;
