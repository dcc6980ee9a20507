// Pins the parity of this repository against the REAL reference ("parity unpinned" in DESIGN.md 2).
//
// Not compiled in this repository (there is no Rust toolchain in its build image).  A maintainer
// of ArielElb/zkSnark-FinalProject copies this file to `src/bin/dump_vectors.rs` of the reference
// (its crate already depends on every arkworks crate used here) and runs
//
//     cargo run --release --bin dump_vectors -- <out_dir>
//
// It writes `ark_dump_*.json` files: inputs AND outputs of arkworks 0.4 at the exact seams this
// library replaces -- Radix2EvaluationDomain fft/ifft (plain and coset), VariableBaseMSM::msm_bigint
// in G1 and G2, LibsnarkReduction::witness_map_from_matrices and
// Groth16::create_proof_with_reduction -- for the reference's own test circuits.  Drop the files
// into `tests/golden/`; `tests/test_arkworks_dump.py` then checks the CPU oracle (`-m "not gpu"`)
// and the CUDA path (`-m gpu`) against them byte for byte.  `tests/golden/make_dump_like.py` writes
// the same format from the repository's oracle, which is what keeps the loader honest until a real
// dump exists.
//
// Format (all field elements as decimal strings of the canonical integer; all points as hex of
// `serialize_compressed`):
//   { "producer": "arkworks", "case": ...,
//     "ntt":  { "log_n", "input", "fft", "ifft", "coset_fft", "coset_ifft" },
//     "msm":  { "scalars", "bases_g1", "g1_result", "bases_g2", "g2_result" },
//     "groth16": { "num_instance", "num_witness", "num_constraints", "a", "b", "c" (rows of [coeff, column]),
//                  "z", "r", "s", "h", "pk": { alpha_g1, beta_g1, delta_g1, beta_g2, gamma_g2, delta_g2,
//                  gamma_abc_g1, a_query, b_g1_query, b_g2_query, h_query, l_query }, "proof" } }
use ark_bls12_381::{Bls12_381, Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use ark_ec::{CurveGroup, Group, VariableBaseMSM};
use ark_ff::{FftField, PrimeField, UniformRand};
use ark_groth16::r1cs_to_qap::{LibsnarkReduction, R1CSToQAP};
use ark_groth16::Groth16;
use ark_poly::{EvaluationDomain, GeneralEvaluationDomain, Radix2EvaluationDomain};
use ark_relations::r1cs::{ConstraintSynthesizer, ConstraintSystem, OptimizationGoal};
use ark_serialize::CanonicalSerialize;
use ark_snark::SNARK;
use ark_std::rand::{rngs::StdRng, SeedableRng};
use ark_r1cs_std::alloc::AllocVar;
use prime_snarks::arkworks::constraints::fibbonaci::FibonacciCircuit;
use prime_snarks::arkworks::matrix_proof_of_work::alloc::FpVar2DVec;
use prime_snarks::arkworks::matrix_proof_of_work::constraints::{matrix_mul, MatrixCircuit};
use prime_snarks::arkworks::matrix_proof_of_work::hasher::hasher;
use std::fmt::Write as _;

fn fr(x: &Fr) -> String { format!("\"{}\"", x.into_bigint()) }
fn frs(v: &[Fr]) -> String { format!("[{}]", v.iter().map(fr).collect::<Vec<_>>().join(",")) }
fn hex<T: CanonicalSerialize>(p: &T) -> String {
    let mut b = Vec::new();
    p.serialize_compressed(&mut b).unwrap();
    let mut s = String::from("\"");
    for x in b { write!(s, "{:02x}", x).unwrap(); }
    s.push('"');
    s
}
fn hexes<T: CanonicalSerialize>(v: &[T]) -> String { format!("[{}]", v.iter().map(hex).collect::<Vec<_>>().join(",")) }
fn rows(m: &[Vec<(Fr, usize)>]) -> String {
    let r: Vec<String> = m.iter().map(|row| {
        format!("[{}]", row.iter().map(|(c, j)| format!("[{},{}]", fr(c), j)).collect::<Vec<_>>().join(","))
    }).collect();
    format!("[{}]", r.join(","))
}

fn ntt_section(rng: &mut StdRng) -> String {
    let log_n = 4u32;
    let n = 1usize << log_n;
    let input: Vec<Fr> = (0..n).map(|_| Fr::rand(rng)).collect();
    let d = Radix2EvaluationDomain::<Fr>::new(n).unwrap();
    let c = d.get_coset(Fr::GENERATOR).unwrap();
    format!("{{\"log_n\":{},\"input\":{},\"fft\":{},\"ifft\":{},\"coset_fft\":{},\"coset_ifft\":{}}}",
            log_n, frs(&input), frs(&d.fft(&input)), frs(&d.ifft(&input)), frs(&c.fft(&input)), frs(&c.ifft(&input)))
}

fn msm_section(rng: &mut StdRng) -> String {
    let n = 40usize;                                  // >= 32: arkworks' window rule leaves the c = 3 branch
    let mut scalars: Vec<Fr> = (0..n).map(|_| Fr::rand(rng)).collect();
    scalars[0] = Fr::from(0u64); scalars[1] = Fr::from(1u64); scalars[2] = -Fr::from(1u64); scalars[3] = Fr::from(65535u64);
    let big: Vec<_> = scalars.iter().map(|s| s.into_bigint()).collect();
    let g1: Vec<G1Affine> = (0..n).map(|_| (G1Projective::generator() * Fr::rand(rng)).into_affine()).collect();
    let g2: Vec<G2Affine> = (0..n).map(|_| (G2Projective::generator() * Fr::rand(rng)).into_affine()).collect();
    let r1 = G1Projective::msm_bigint(&g1, &big).into_affine();
    let r2 = G2Projective::msm_bigint(&g2, &big).into_affine();
    format!("{{\"scalars\":{},\"bases_g1\":{},\"g1_result\":{},\"bases_g2\":{},\"g2_result\":{}}}",
            frs(&scalars), hexes(&g1), hex(&r1), hexes(&g2), hex(&r2))
}

fn groth16_section<C: ConstraintSynthesizer<Fr> + Clone>(circuit: C, rng: &mut StdRng) -> String {
    let (pk, vk) = Groth16::<Bls12_381>::circuit_specific_setup(circuit.clone(), rng).unwrap();
    // synthesis exactly as ark-groth16's create_proof_with_reduction does it
    let cs = ConstraintSystem::<Fr>::new_ref();
    cs.set_optimization_goal(OptimizationGoal::Constraints);
    circuit.clone().generate_constraints(cs.clone()).unwrap();
    cs.finalize();
    let m = cs.to_matrices().unwrap();
    let z: Vec<Fr> = {
        let p = cs.borrow().unwrap();
        [&p.instance_assignment[..], &p.witness_assignment[..]].concat()
    };
    let r = Fr::rand(rng);
    let s = Fr::rand(rng);
    let proof = Groth16::<Bls12_381>::create_proof_with_reduction(circuit, &pk, r, s).unwrap();
    let h = LibsnarkReduction::witness_map_from_matrices::<Fr, GeneralEvaluationDomain<Fr>>(
        &m, m.num_instance_variables, m.num_constraints, &z).unwrap();
    let pk_json = format!(
        "{{\"alpha_g1\":{},\"beta_g1\":{},\"delta_g1\":{},\"beta_g2\":{},\"gamma_g2\":{},\"delta_g2\":{},\"gamma_abc_g1\":{},\
         \"a_query\":{},\"b_g1_query\":{},\"b_g2_query\":{},\"h_query\":{},\"l_query\":{}}}",
        hex(&vk.alpha_g1), hex(&pk.beta_g1), hex(&pk.delta_g1), hex(&vk.beta_g2), hex(&vk.gamma_g2), hex(&vk.delta_g2),
        hexes(&vk.gamma_abc_g1), hexes(&pk.a_query), hexes(&pk.b_g1_query), hexes(&pk.b_g2_query), hexes(&pk.h_query),
        hexes(&pk.l_query));
    format!("{{\"num_instance\":{},\"num_witness\":{},\"num_constraints\":{},\"a\":{},\"b\":{},\"c\":{},\"z\":{},\"r\":{},\"s\":{},\
             \"h\":{},\"pk\":{},\"proof\":{}}}",
            m.num_instance_variables, m.num_witness_variables, m.num_constraints, rows(&m.a), rows(&m.b), rows(&m.c),
            frs(&z), fr(&r), fr(&s), frs(&h), pk_json, hex(&proof))
}

fn write_case<C: ConstraintSynthesizer<Fr> + Clone>(dir: &str, case: &str, circuit: C, seed: u64) {
    let mut rng = StdRng::seed_from_u64(seed);
    let body = format!("{{\"producer\":\"arkworks\",\"case\":\"{}\",\"ntt\":{},\"msm\":{},\"groth16\":{}}}\n",
                       case, ntt_section(&mut rng), msm_section(&mut rng), groth16_section(circuit, &mut rng));
    let path = format!("{}/ark_dump_{}.json", dir, case);
    std::fs::write(&path, body).unwrap();
    println!("wrote {}", path);
}

fn main() {
    let dir = std::env::args().nth(1).unwrap_or_else(|| ".".to_string());
    // src/arkworks/constraints/fibbonaci.rs:196-198 (a = 0, b = 1, 10 steps -> 55) and the handler's 1000 steps
    write_case(&dir, "fibonacci_0_1_10",
               FibonacciCircuit::<Fr> { a: Some(Fr::from(0u64)), b: Some(Fr::from(1u64)), num_of_steps: 10, result: Some(Fr::from(55u64)) },
               0xB2000004);
    // matrix_proof_of_work/constraints.rs:231-271: [[1,2],[3,4]] * [[4,3],[2,1]], hashes computed the way the
    // reference's own test computes them (a scratch constraint system just to obtain the three digests)
    for (case, a, b) in [
        ("matrix_2x2", vec![vec![1u64, 2], vec![3, 4]], vec![vec![4u64, 3], vec![2, 1]]),
        ("matrix_4x4_ones", vec![vec![1u64; 4]; 4], vec![vec![1u64; 4]; 4]),
    ] {
        let scratch = ConstraintSystem::<Fr>::new_ref();
        let av = FpVar2DVec::new_witness(scratch.clone(), || Ok(a.clone())).unwrap();
        let bv = FpVar2DVec::new_witness(scratch.clone(), || Ok(b.clone())).unwrap();
        let cv = matrix_mul(scratch.clone(), av.clone(), bv.clone());
        let (ha, hb, hc) = (hasher(&av).unwrap()[0], hasher(&bv).unwrap()[0], hasher(&cv).unwrap()[0]);
        write_case(&dir, case, MatrixCircuit::<Fr>::new(a, b, ha, hb, hc), 0xB2000004);
    }
    // The prime circuit (prime_snark/prime_circut.rs:361, x = 5) is dumped the same way: build the circuit value
    // exactly as its own test does and call write_case(&dir, "prime_5", circuit, 0xB2000004).  The matrices travel
    // inside the dump, so the checker needs no knowledge of the Poseidon / SHA-256 gadgets.
}
