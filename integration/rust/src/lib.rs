//! `verify_proofs_batch` for ChainSafe/halo2-verifier on B200: the reference's verification surface
//! (`ParamsKZG`, `VerifyingKey`, `verify_proof`, halo2_verifier/src/lib.rs:29-46) plus a batch entry point that
//! packs the serialized proofs and public inputs and calls the `extern "C"` layer of libh2v_b200 (include/h2v.h).
//! There is no CPU fallback on this path: without a CUDA device `BatchVerifier::new` fails.
#![cfg(feature = "cuda")]

pub mod cuda;
pub use cuda::{BatchVerifier, MultiOpen, ShardChannel, TranscriptHash};

use halo2_verifier::{
    halo2curves::bn256::{Fr, G1Affine},
    plonk::Error,
    ParamsKZG, VerifyingKey,
};

/// The batch counterpart of `verify_proof` (lib.rs:33-46) with `AccumulatorStrategy` semantics
/// (poly/kzg/strategy.rs:125-140): one folded pairing check for all proofs, and - when that rejects - the per-proof
/// re-check the reference prescribes (poly/strategy.rs:26-30).  One entry per proof: `Ok(())` or the `plonk::Error`
/// `verify_proof` would have returned for that proof.
pub fn verify_proofs_batch(
    params: &ParamsKZG,
    vk: &VerifyingKey<G1Affine>,
    proofs: &[&[u8]],
    instances: &[&[&[Fr]]],
) -> Result<Vec<Result<(), Error>>, String> {
    let mut bv = BatchVerifier::new(params, vk, MultiOpen::Shplonk, TranscriptHash::Blake2b, 0)?;
    bv.verify_proofs_batch(proofs, instances)
}
