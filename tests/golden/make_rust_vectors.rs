// make_rust_vectors.rs — dumps golden vectors FROM THE REFERENCE ITSELF (the Rust crate AHartNtkn/odesat).
//
// This image has no cargo/rustc, so this file has never been compiled here.  A maintainer with a Rust
// toolchain runs it once inside the reference crate; the JSON it prints pins this repo's CPU oracle
// (tests/test_rust_vectors.py consumes tests/golden/rust_*.json when present):
//
//     cp make_rust_vectors.rs  <odesat>/examples/make_rust_vectors.rs
//     cd <odesat>
//     cargo run --release --example make_rust_vectors -- tests/easy.cnf  > rust_easy.json
//     cargo run --release --example make_rust_vectors -- tests/small.cnf > rust_small.json
//     cp rust_easy.json rust_small.json  <odesat_b200>/tests/golden/
//
// Only the crate's externally callable API is used: `SlabState` has private fields and no constructor, so
// `compute_derivatives` / `euler_step*` cannot be called from outside `odesat::system`; `simulate` with
// `steps = Some(1)` is exactly one `euler_step_fixed` / `euler_step` (src/system.rs:190-196, 206-220), which is what
// is dumped.  Every f64 is written as the hex of its IEEE-754 bit pattern (exact, no formatting ambiguity); the
// normalised formula is written as CSR because `normalize_cnf_variables` numbers the variables in HashSet iteration
// order (src/cnf.rs:206-219), which differs from run to run.  No dependency beyond the crate's own (ndarray).
use ndarray::prelude::*;
use odesat::cnf::*;
use odesat::system::*;
use std::env;
use std::fs;

// Deterministic v0 ∈ [-1, 1): SplitMix64, top 53 bits (the crate itself uses an OS-seeded thread_rng).
fn splitmix(state: &mut u64) -> u64 {
    *state = state.wrapping_add(0x9E3779B97F4A7C15);
    let mut z = *state;
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^ (z >> 31)
}

fn v0(n: usize, seed: u64) -> Array1<f64> {
    let mut s = seed;
    Array1::from_iter((0..n).map(|_| (splitmix(&mut s) >> 11) as f64 * (1.0 / 9007199254740992.0) * 2.0 - 1.0))
}

fn hex(a: &Array1<f64>) -> String {
    let items: Vec<String> = a.iter().map(|x| format!("\"{:016x}\"", x.to_bits())).collect();
    format!("[{}]", items.join(","))
}

fn state_json(s: &State) -> String {
    format!("{{\"v\":{},\"xs\":{},\"xl\":{}}}", hex(&s.v), hex(&s.xs), hex(&s.xl))
}

fn bools(b: &[bool]) -> String {
    let items: Vec<&str> = b.iter().map(|&x| if x { "1" } else { "0" }).collect();
    format!("[{}]", items.join(","))
}

fn fresh(formula: &CNFFormula, normalized: &CNFFormula, seed: u64) -> State {
    State {
        v: v0(normalized.varnum, seed),
        xs: init_short_term_memory(formula), // as src/main.rs:172 does (the un-normalised formula, same result)
        xl: Array1::ones(normalized.clauses.len()),
    }
}

fn main() {
    let path = env::args().nth(1).expect("usage: make_rust_vectors <file.cnf>");
    let text = fs::read_to_string(&path).expect("read");
    let formula = parse_dimacs_format(&text);
    let (_map, nf) = normalize_cnf_variables(&formula);

    // normalised formula as CSR: lits = ±(index + 1)
    let mut off: Vec<usize> = vec![0];
    let mut lits: Vec<i64> = vec![];
    for c in nf.clauses.iter() {
        for l in c.literals.iter() {
            let x = (l.variable as i64) + 1;
            lits.push(if l.is_negated { -x } else { x });
        }
        off.push(lits.len());
    }
    let offs: Vec<String> = off.iter().map(|x| x.to_string()).collect();
    let litss: Vec<String> = lits.iter().map(|x| x.to_string()).collect();

    let mut out = String::new();
    out.push_str(&format!(
        "{{\"source\":\"{}\",\"varnum\":{},\"clause_off\":[{}],\"lits\":[{}],\"cases\":[",
        path,
        nf.varnum,
        offs.join(","),
        litss.join(",")
    ));

    let mut first = true;
    let mut emit = |out: &mut String, name: &str, body: String| {
        if !first {
            out.push(',');
        }
        first = false;
        out.push_str(&format!("{{\"name\":\"{}\",{}}}", name, body));
    };

    // fixed steps (dt = 0.01): 1, 2 and 100 calls of euler_step_fixed from the same start
    for &(steps, seed) in &[(1usize, 11u64), (2, 11), (100, 11), (100, 12)] {
        let mut s = fresh(&formula, &nf, seed);
        let start = state_json(&s);
        let res = simulate(&mut s, &nf, None, Some(0.01), Some(steps), None);
        emit(
            &mut out,
            "fixed",
            format!(
                "\"seed\":{},\"steps\":{},\"step_size\":\"{:016x}\",\"start\":{},\"end\":{},\"assignment\":{}",
                seed,
                steps,
                0.01f64.to_bits(),
                start,
                state_json(&s),
                bools(&res)
            ),
        );
    }
    // adaptive steps (tolerance = 1e-3 default, dt0 = 0.01): 1, 2, 3 and 50 calls of euler_step
    for &(steps, seed) in &[(1usize, 21u64), (2, 21), (3, 21), (50, 21), (50, 22)] {
        let mut s = fresh(&formula, &nf, seed);
        let start = state_json(&s);
        let res = simulate(&mut s, &nf, None, None, Some(steps), None);
        emit(
            &mut out,
            "adaptive",
            format!(
                "\"seed\":{},\"steps\":{},\"start\":{},\"end\":{},\"assignment\":{}",
                seed,
                steps,
                start,
                state_json(&s),
                bools(&res)
            ),
        );
    }
    // simulate_inter: 4 replicas, fixed and adaptive (shared dt, src/system.rs:314), 40 outer steps
    for &adaptive in &[false, true] {
        let mut states: Vec<State> = (0..4).map(|r| fresh(&formula, &nf, 31 + r as u64)).collect();
        let start: Vec<String> = states.iter().map(state_json).collect();
        let step = if adaptive { None } else { Some(0.01) };
        let res = simulate_inter(&mut states, &nf, None, step, Some(40), None);
        let end: Vec<String> = states.iter().map(state_json).collect();
        emit(
            &mut out,
            "inter",
            format!(
                "\"adaptive\":{},\"steps\":40,\"seeds\":[31,32,33,34],\"start\":[{}],\"end\":[{}],\"assignment\":{}",
                adaptive,
                start.join(","),
                end.join(","),
                bools(&res)
            ),
        );
    }
    // max_error and update_state on two states
    {
        let a = fresh(&formula, &nf, 41);
        let mut b = fresh(&formula, &nf, 42);
        let err = max_error(&a, &b);
        let before = state_json(&b);
        update_state(&mut b, &a, 0.25, nf.clauses.len()); // `a` used as a derivative container (src/system.rs:93)
        emit(
            &mut out,
            "max_error_update",
            format!(
                "\"a\":{},\"b\":{},\"max_error\":\"{:016x}\",\"dt\":\"{:016x}\",\"b_after_update_with_a\":{}",
                state_json(&a),
                before,
                err.to_bits(),
                0.25f64.to_bits(),
                state_json(&b)
            ),
        );
    }
    out.push_str("]}");
    println!("{}", out);
}
