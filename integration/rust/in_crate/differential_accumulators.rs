//! In-crate differential test for the accumulators.  The reference keeps `DualMSM::{left, right}` and
//! `GuardKZG::msm_accumulator` `pub(crate)` (poly/kzg/msm.rs:148-156, poly/kzg/strategy.rs:23-31), so the per-proof
//! accumulated G1 points can only be read from INSIDE the `halo2_verifier` crate.  To run:
//!
//!   1. copy this file to `halo2_verifier/src/differential_accumulators.rs`
//!   2. add to `halo2_verifier/src/lib.rs`:   `#[cfg(test)] mod differential_accumulators;`
//!   3. add dev-dependencies `serde_json = "1"`, `hex = "0.4"` (serde_json and hex-literal are already listed)
//!   4. `H2V_B200_DIR=/path/to/this/repo cargo test -p halo2_verifier differential_accumulators -- --nocapture`
//!
//! It replays every golden proof through `V::verify_proof` with a strategy that keeps the guard, evaluates both MSM
//! channels exactly as `DualMSM::check` does (msm.rs:189-190) and compares the affine (L_j, R_j) with the golden bytes
//! (x | y, 32-byte little-endian canonical each, all-zero = identity), then folds them with the golden r_i in the
//! AccumulatorStrategy convention (strategy.rs:125-136: scale BEFORE each proof) and compares the folded (L, R).
#![cfg(test)]
use crate::{
    helpers::SerdeFormat,
    plonk::Error,
    poly::{
        commitment::{Verifier, MSM},
        kzg::{
            commitment::{KZGCommitmentScheme, ParamsKZG},
            msm::DualMSM,
            multiopen::{VerifierGWC, VerifierSHPLONK},
            strategy::GuardKZG,
        },
        strategy::VerificationStrategy,
    },
    transcript::{Blake2bRead, Challenge255, Keccak256Read, TranscriptReadBuffer},
    verify_proof, VerifyingKey,
};
use ff::PrimeField;
use group::{prime::PrimeCurveAffine, Curve};
use halo2curves::{
    bn256::{Bn256, Fr, G1Affine},
    CurveAffine,
};

/// hands the fresh accumulator to the multi-open verifier and returns the filled one instead of checking it
struct Keep<'p>(&'p ParamsKZG<Bn256>);
impl<'p, V: Verifier<'p, KZGCommitmentScheme<Bn256>, MSMAccumulator = DualMSM<'p, Bn256>, Guard = GuardKZG<'p, Bn256>>>
    VerificationStrategy<'p, KZGCommitmentScheme<Bn256>, V> for Keep<'p>
{
    type Output = DualMSM<'p, Bn256>;
    fn new(params: &'p ParamsKZG<Bn256>) -> Self {
        Keep(params)
    }
    fn process(self, f: impl FnOnce(V::MSMAccumulator) -> Result<V::Guard, Error>) -> Result<Self::Output, Error> {
        Ok(f(DualMSM::new(self.0))?.msm_accumulator)
    }
    fn finalize(self) -> bool {
        unreachable!()
    }
}

fn point_bytes(p: G1Affine) -> Vec<u8> {
    if bool::from(p.is_identity()) {
        return vec![0u8; 64];
    }
    let c = p.coordinates().unwrap();
    [c.x().to_repr().as_ref(), c.y().to_repr().as_ref()].concat()
}
fn fr_from_hex(s: &str) -> Fr {
    let mut be = hex::decode(format!("{:0>64}", s.trim_start_matches("0x"))).unwrap();
    be.reverse();
    let mut repr = <Fr as PrimeField>::Repr::default();
    repr.as_mut().copy_from_slice(&be);
    Option::from(Fr::from_repr(repr)).unwrap()
}

#[test]
fn accumulators_equal_golden() {
    let root = std::env::var("H2V_B200_DIR").expect("H2V_B200_DIR");
    for name in ["vm_k8_shplonk_blake2b", "vm_k8_gwc_keccak", "sh_k8_shplonk_keccak", "mix_k6_shplonk_blake2b", "mix_k6_gwc_blake2b"] {
        let g: serde_json::Value = serde_json::from_slice(&std::fs::read(format!("{root}/tests/golden/{name}.json")).unwrap()).unwrap();
        let (mo, hash) = (g["multiopen"].as_str().unwrap(), g["hash"].as_str().unwrap());
        let pbytes = hex::decode(g["params"].as_str().unwrap()).unwrap();
        let vbytes = hex::decode(g["vk"].as_str().unwrap()).unwrap();
        let fmt = if g["vk_format"].as_u64().unwrap() == 0 { SerdeFormat::Processed } else { SerdeFormat::RawBytes };
        let params = ParamsKZG::<Bn256>::read(&mut &pbytes[..]).unwrap();
        let vk = VerifyingKey::<G1Affine>::read(&mut &vbytes[..], fmt).unwrap();
        let rs: Vec<Fr> = g["rlc_scalars"].as_array().unwrap().iter().map(|v| fr_from_hex(v.as_str().unwrap())).collect();
        let mut fold = DualMSM::new(&params);
        for (j, p) in g["proofs"].as_array().unwrap().iter().enumerate() {
            let proof = hex::decode(p["proof"].as_str().unwrap()).unwrap();
            let inst: Vec<Vec<Fr>> = p["instances"].as_array().unwrap().iter()
                .map(|c| c.as_array().unwrap().iter().map(|v| fr_from_hex(v.as_str().unwrap())).collect()).collect();
            let cols: Vec<&[Fr]> = inst.iter().map(|c| &c[..]).collect();
            let instances: [&[&[Fr]]; 1] = [&cols[..]];
            macro_rules! run {
                ($tr:expr, $v:ty) => {{
                    let mut t = $tr;
                    verify_proof::<KZGCommitmentScheme<Bn256>, $v, Challenge255<G1Affine>, _, Keep>(&params, &vk, Keep(&params), &instances, &mut t)
                }};
            }
            let acc = match (mo, hash) {
                ("shplonk", "blake2b") => run!(Blake2bRead::<_, G1Affine, Challenge255<_>>::init(&proof[..]), VerifierSHPLONK<Bn256>),
                ("shplonk", _) => run!(Keccak256Read::<_, G1Affine, Challenge255<_>>::init(&proof[..]), VerifierSHPLONK<Bn256>),
                ("gwc", "blake2b") => run!(Blake2bRead::<_, G1Affine, Challenge255<_>>::init(&proof[..]), VerifierGWC<Bn256>),
                _ => run!(Keccak256Read::<_, G1Affine, Challenge255<_>>::init(&proof[..]), VerifierGWC<Bn256>),
            };
            fold.scale(rs[j]); // AccumulatorStrategy::process scales the accumulator before every proof (strategy.rs:129)
            match (acc, p.get("accum")) {
                (Ok(d), Some(want)) => {
                    let got = [point_bytes(d.left.eval().to_affine()), point_bytes(d.right.eval().to_affine())].concat();
                    assert_eq!(hex::encode(got), want.as_str().unwrap(), "{name} proof {j}: (L_j, R_j)");
                    fold.add_msm(d);
                }
                (Err(_), None) => {}
                (a, w) => panic!("{name} proof {j}: reference produced accumulators = {}, golden has them = {}", a.is_ok(), w.is_some()),
            }
        }
        let got = [point_bytes(fold.left.eval().to_affine()), point_bytes(fold.right.eval().to_affine())].concat();
        assert_eq!(hex::encode(got), g["folded"].as_str().unwrap(), "{name}: folded (L, R)");
        assert_eq!(fold.check(), g["folded_ok"].as_bool().unwrap(), "{name}: batch verdict");
    }
}
