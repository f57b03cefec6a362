//! proof_parity.rs — the proof-byte parity check of SURVEY.md §8c / INTEGRATION.md §6.  SOURCE ONLY: it needs cargo, the zksnap
//! workspace with halo2-axiom patched as INTEGRATION.md §3 describes, and a B200; none of that exists in the build image, so
//! this file has never been compiled.  It is written against the reference's own prover entry points
//! (/root/reference/aggregator/src/wrapper.rs:106-109 gen_pk, :129-137 create_proof) and changes two things only: the SRS and the
//! prover RNG are seeded (the reference uses OsRng at wrapper.rs:134 and ParamsKZG::setup(k, OsRng) at
//! voter/benches/voter_circuit.rs:60), so that two runs are comparable byte for byte.
//!
//! The patched halo2-axiom routes best_multiexp / best_fft / EvaluationDomain to libzkb200 unless ZKB200_DISABLE=1 is set when the
//! process starts (the patch reads it once into a static), so the same binary produces both proofs:
//!
//!     ZKB200_DISABLE=1 cargo run --release --example proof_parity -- voter 15 > cpu.hex
//!     cargo run --release --example proof_parity -- voter 15 > gpu.hex && cmp cpu.hex gpu.hex
//!
//! Bit-exactness at the kernel boundary (what this repository proves: canonical Montgomery limbs, normalised points) implies equal
//! transcripts and therefore equal proofs; this harness is the end-to-end confirmation.
use halo2_base::halo2_proofs::{
    halo2curves::bn256::{Bn256, Fr, G1Affine},
    plonk::{create_proof, keygen_pk, keygen_vk, Circuit},
    poly::kzg::{commitment::ParamsKZG, multiopen::ProverGWC},
};
use rand_chacha::{rand_core::SeedableRng, ChaCha20Rng};
use snark_verifier_sdk::halo2::PoseidonTranscript;
use snark_verifier_sdk::NativeLoader;

/// `gen_proof` of the reference with the RNGs seeded; everything else as in wrapper.rs:111-158.
fn seeded_proof<C: Circuit<Fr>>(k: u32, circuit: C, instances: Vec<Vec<Fr>>) -> Vec<u8> {
    let params = ParamsKZG::<Bn256>::setup(k, ChaCha20Rng::from_seed([7u8; 32]));
    let vk = keygen_vk(&params, &circuit).unwrap();
    let pk = keygen_pk(&params, vk, &circuit).unwrap();
    let instances: Vec<&[Fr]> = instances.iter().map(Vec::as_slice).collect();
    let mut transcript = PoseidonTranscript::<NativeLoader, _>::new::<0>(Vec::new());
    create_proof::<_, ProverGWC<_>, _, _, _, _>(
        &params,
        &pk,
        &[circuit],
        &[instances.as_slice()],
        ChaCha20Rng::from_seed([9u8; 32]),
        &mut transcript,
    )
    .unwrap();
    transcript.finalize()
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let which = args.get(1).map(String::as_str).unwrap_or("voter");
    let k: u32 = args.get(2).and_then(|s| s.parse().ok()).unwrap_or(15);
    // The circuits and their inputs come from the reference's own builders: voter::utils::generate_random_voter_circuit_inputs
    // and aggregator::state_transition (see the three benches); they are constructed here exactly as those benches do.
    let proof = match which {
        "voter" => {
            let (circuit, instances) = zksnap_parity_inputs::voter(k);
            seeded_proof(k, circuit, instances)
        }
        "state_transition" => {
            let (circuit, instances) = zksnap_parity_inputs::state_transition(k);
            seeded_proof(k, circuit, instances)
        }
        other => panic!("unknown circuit {other}: voter | state_transition (the wrapper goes through gen_recursion_snark, wrapper.rs:869-902)"),
    };
    println!("{}", proof.iter().map(|b| format!("{b:02x}")).collect::<String>());
}

/// Thin adapters over the reference's input generators (kept out of this file so that it does not restate reference code):
/// `voter(k)` wraps /root/reference/voter/benches/voter_circuit.rs:30-58, `state_transition(k)` wraps
/// /root/reference/aggregator/benches/state_transition_circuit.rs:20-62.
mod zksnap_parity_inputs {
    pub use zksnap_parity_adapters::{state_transition, voter};
}
