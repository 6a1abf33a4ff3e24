//! Runs the UNMODIFIED reference solver on a scene file and prints the final state as f32 bit patterns.
//!
//!   ref-harness scene.txt > state.txt
//!
//! Scene file (written by export_inputs.py, one record per line, floats as 8 hex digits of their bit pattern so
//! nothing is lost to decimal formatting):
//!   bounds  x y w h
//!   gravity x y
//!   particle x y
//!   plink a b len                 (ParticleLink, insertion order)
//!   circle x y r
//!   clink a b len                 (CircleLink)
//!   polygon is_static n x0 y0 x1 y1 ...     (Polygon::new: perimeter links, polygon.rs:84-123)
//!   run n_updates dt
//! Output: `p x y px py` per particle, `c x y px py` per circle, `g k x y px py` per polygon point, `gc k cx cy`.
use bendy2d::circle::Circle;
use bendy2d::link::{CircleLink, Link, ParticleLink};
use bendy2d::particle::Particle;
use bendy2d::polygon::Polygon;
use bendy2d::solver::{Bounds, Solver};
use nalgebra::Vector2;
use std::io::{BufRead, BufReader};

fn f(tok: &str) -> f32 {
    f32::from_bits(u32::from_str_radix(tok, 16).expect("hex float"))
}
fn h(v: f32) -> String {
    format!("{:08x}", v.to_bits())
}

fn main() {
    let path = std::env::args().nth(1).expect("usage: ref-harness scene.txt");
    let file = BufReader::new(std::fs::File::open(path).expect("open scene"));
    let mut s = Solver::new();
    let (mut n_updates, mut dt) = (0usize, 0f32);
    for line in file.lines() {
        let line = line.unwrap();
        let t: Vec<&str> = line.split_whitespace().collect();
        if t.is_empty() {
            continue;
        }
        match t[0] {
            "bounds" => s.bounds = Bounds { pos: Vector2::new(f(t[1]), f(t[2])), size: Vector2::new(f(t[3]), f(t[4])) },
            "gravity" => s.gravity = Vector2::new(f(t[1]), f(t[2])),
            "particle" => s.add_particle(Vector2::new(f(t[1]), f(t[2]))),
            "plink" => s.add_particle_link(ParticleLink {
                link: Link { particle_a: t[1].parse().unwrap(), particle_b: t[2].parse().unwrap(), target_distance: f(t[3]) },
            }),
            "circle" => s.add_circle(Circle { point: Particle::new(Vector2::new(f(t[1]), f(t[2]))), radius: f(t[3]) }),
            "clink" => s.add_circle_link(CircleLink {
                link: Link { particle_a: t[1].parse().unwrap(), particle_b: t[2].parse().unwrap(), target_distance: f(t[3]) },
            }),
            "polygon" => {
                let is_static = t[1] == "1";
                let n: usize = t[2].parse().unwrap();
                let pts: Vec<Vector2<f32>> = (0..n).map(|i| Vector2::new(f(t[3 + 2 * i]), f(t[4 + 2 * i]))).collect();
                s.add_polygon(Polygon::new(pts, is_static));
            }
            "run" => {
                n_updates = t[1].parse().unwrap();
                dt = f(t[2]);
            }
            other => panic!("unknown record {other}"),
        }
    }
    for _ in 0..n_updates {
        s.update(dt); // solver.rs:106-116; sub_steps is fixed at 1 in the reference
    }
    for p in s.get_particles() {
        println!("p {} {} {} {}", h(p.pos.x), h(p.pos.y), h(p.prev_pos.x), h(p.prev_pos.y));
    }
    for c in s.get_circles() {
        println!("c {} {} {} {}", h(c.point.pos.x), h(c.point.pos.y), h(c.point.prev_pos.x), h(c.point.prev_pos.y));
    }
    for (k, g) in s.get_polygons().iter().enumerate() {
        for p in &g.particles {
            println!("g {} {} {} {} {}", k, h(p.pos.x), h(p.pos.y), h(p.prev_pos.x), h(p.prev_pos.y));
        }
        println!("gc {} {} {}", k, h(g.center.x), h(g.center.y));
    }
}
