//! Differential test: the REAL reference (`halo2_verifier::verify_proof`) against this repository's committed golden
//! vectors (tests/golden/*.json, produced by the CPU oracle) and against the CUDA path, on the same bytes.
//!
//! What it pins (SURVEY.md 8c "parity unpinned"): per proof the accept / reject class, every transcript challenge in
//! squeeze order (through a recording transcript), and - CUDA vs golden - the per-proof accumulators (L_j, R_j) and the
//! folded (L, R).  The reference keeps its accumulators `pub(crate)` (poly/kzg/msm.rs:148-156), so comparing THEM with
//! the golden values needs the in-crate variant `integration/rust/in_crate/differential_accumulators.rs`.
//!
//!     H2V_B200_DIR=/path/to/this/repo cargo test --release -- --nocapture
//!
//! NOT RUN in this repository's CI (no Rust toolchain in the build image).  Until it has been run by someone with cargo,
//! DESIGN.md section 2 keeps the parity status at "bit-exact against our restatement of the reference".
use ff::PrimeField;
use halo2_verifier::{
    halo2curves::{
        bn256::{Bn256, Fr, G1Affine},
        CurveAffine,
    },
    helpers::SerdeFormat,
    io,
    plonk::Error,
    poly::kzg::{
        commitment::KZGCommitmentScheme,
        multiopen::{VerifierGWC, VerifierSHPLONK},
        strategy::SingleStrategy,
    },
    transcript::{Blake2bRead, Challenge255, EncodedChallenge, Keccak256Read, Transcript, TranscriptRead, TranscriptReadBuffer},
    verify_proof, ParamsKZG, VerifyingKey,
};
use halo2_verifier_cuda::{BatchVerifier, MultiOpen, TranscriptHash};
use serde_json::Value;

/// Wraps a reference transcript and records every squeezed challenge as a scalar (transcript/mod.rs:42-89).
struct Recording<T> {
    inner: T,
    challenges: Vec<Fr>,
}
impl<T: Transcript<G1Affine, Challenge255<G1Affine>>> Transcript<G1Affine, Challenge255<G1Affine>> for Recording<T> {
    fn squeeze_challenge(&mut self) -> Challenge255<G1Affine> {
        let c = self.inner.squeeze_challenge();
        self.challenges.push(c.get_scalar());
        c
    }
    fn common_point(&mut self, point: G1Affine) -> io::Result<()> {
        self.inner.common_point(point)
    }
    fn common_scalar(&mut self, scalar: Fr) -> io::Result<()> {
        self.inner.common_scalar(scalar)
    }
}
impl<T: TranscriptRead<G1Affine, Challenge255<G1Affine>>> TranscriptRead<G1Affine, Challenge255<G1Affine>> for Recording<T> {
    fn read_point(&mut self) -> io::Result<G1Affine> {
        self.inner.read_point()
    }
    fn read_scalar(&mut self) -> io::Result<Fr> {
        self.inner.read_scalar()
    }
}

fn status_of(r: &Result<(), Error>) -> u8 {
    match r {
        Ok(()) => 0,
        Err(Error::InvalidInstances) => 1,
        Err(Error::Transcript(_)) => 2,
        Err(Error::Opening) => 3,
        Err(Error::ConstraintSystemFailure) => 4,
        Err(_) => 255,
    }
}

fn fr_from_hex(s: &str) -> Fr {
    let s = s.trim_start_matches("0x");
    let mut be = hex::decode(format!("{:0>64}", s)).unwrap();
    be.reverse();
    let mut repr = <Fr as PrimeField>::Repr::default();
    repr.as_mut().copy_from_slice(&be);
    Option::from(Fr::from_repr(repr)).expect("canonical scalar")
}
fn fr_to_le(f: &Fr) -> Vec<u8> {
    f.to_repr().as_ref().to_vec()
}

/// the reference's verify_proof with SingleStrategy on one proof: (result, challenges)
fn reference_verify(params: &ParamsKZG<Bn256>, vk: &VerifyingKey<G1Affine>, proof: &[u8], inst: &[Vec<Fr>], mo: &str, hash: &str) -> (Result<(), Error>, Vec<Fr>) {
    let cols: Vec<&[Fr]> = inst.iter().map(|c| &c[..]).collect();
    let instances: [&[&[Fr]]; 1] = [&cols[..]];
    macro_rules! run {
        ($tr:expr, $v:ty) => {{
            let mut t = Recording { inner: $tr, challenges: vec![] };
            let r = verify_proof::<KZGCommitmentScheme<Bn256>, $v, Challenge255<G1Affine>, _, SingleStrategy<Bn256>>(
                params, vk, SingleStrategy::new(params), &instances, &mut t);
            (r, t.challenges)
        }};
    }
    match (mo, hash) {
        ("shplonk", "blake2b") => run!(Blake2bRead::<_, G1Affine, Challenge255<_>>::init(proof), VerifierSHPLONK<Bn256>),
        ("shplonk", _) => run!(Keccak256Read::<_, G1Affine, Challenge255<_>>::init(proof), VerifierSHPLONK<Bn256>),
        ("gwc", "blake2b") => run!(Blake2bRead::<_, G1Affine, Challenge255<_>>::init(proof), VerifierGWC<Bn256>),
        _ => run!(Keccak256Read::<_, G1Affine, Challenge255<_>>::init(proof), VerifierGWC<Bn256>),
    }
}

#[test]
fn reference_equals_golden_equals_cuda() {
    let root = std::env::var("H2V_B200_DIR").unwrap_or_else(|_| format!("{}/../..", env!("CARGO_MANIFEST_DIR")));
    let mut checked = 0;
    for entry in std::fs::read_dir(format!("{root}/tests/golden")).unwrap() {
        let path = entry.unwrap().path();
        if path.extension().map(|e| e != "json").unwrap_or(true) {
            continue;
        }
        let g: Value = serde_json::from_slice(&std::fs::read(&path).unwrap()).unwrap();
        // single-circuit-instance vectors with params + vk bytes (the honest-prover and srs_kat files have their own layouts)
        if g.get("circuit_instances").is_some() || g.get("proofs").is_none() || g.get("params").is_none() {
            continue;
        }
        let (mo, hash) = (g["multiopen"].as_str().unwrap(), g["hash"].as_str().unwrap());
        let pbytes = hex::decode(g["params"].as_str().unwrap()).unwrap();
        let vbytes = hex::decode(g["vk"].as_str().unwrap()).unwrap();
        let fmt = if g["vk_format"].as_u64().unwrap() == 0 { SerdeFormat::Processed } else { SerdeFormat::RawBytes };
        let params = ParamsKZG::<Bn256>::read(&mut &pbytes[..]).expect("params (Processed form, commitment.rs:133-139)");
        let vk = VerifyingKey::<G1Affine>::read(&mut &vbytes[..], fmt).expect("vk (plonk/vk.rs:76-115)");
        let proofs: Vec<Vec<u8>> = g["proofs"].as_array().unwrap().iter().map(|p| hex::decode(p["proof"].as_str().unwrap()).unwrap()).collect();
        let insts: Vec<Vec<Vec<Fr>>> = g["proofs"].as_array().unwrap().iter()
            .map(|p| p["instances"].as_array().unwrap().iter().map(|c| c.as_array().unwrap().iter().map(|v| fr_from_hex(v.as_str().unwrap())).collect()).collect())
            .collect();
        let rlc: Vec<Fr> = g["rlc_scalars"].as_array().unwrap().iter().map(|v| fr_from_hex(v.as_str().unwrap())).collect();

        // ---- 1. the real reference against the golden vectors
        for (j, p) in g["proofs"].as_array().unwrap().iter().enumerate() {
            let (res, chal) = reference_verify(&params, &vk, &proofs[j], &insts[j], mo, hash);
            assert_eq!(status_of(&res) as u64, p["status"].as_u64().unwrap(), "{path:?} proof {j}: verdict class");
            let want: Vec<Fr> = p["challenges"].as_array().unwrap().iter().map(|v| fr_from_hex(v.as_str().unwrap())).collect();
            // a failing read ends the reference's transcript early: it must agree on every challenge it did squeeze
            assert!(chal.len() <= want.len() && chal[..] == want[..chal.len()], "{path:?} proof {j}: transcript challenges");
            if res.is_ok() {
                assert_eq!(chal.len(), want.len(), "{path:?} proof {j}: number of squeezes");
            }
        }

        // ---- 2. the CUDA path against the golden vectors (statuses, challenges, accumulators, folded (L, R))
        let mut bv = BatchVerifier::new(&params, &vk, if mo == "gwc" { MultiOpen::Gwc } else { MultiOpen::Shplonk },
                                        if hash == "blake2b" { TranscriptHash::Blake2b } else { TranscriptHash::Keccak256 }, 0).unwrap();
        let pr: Vec<&[u8]> = proofs.iter().map(|p| &p[..]).collect();
        let cols: Vec<Vec<&[Fr]>> = insts.iter().map(|i| i.iter().map(|c| &c[..]).collect()).collect();
        let ins: Vec<&[&[Fr]]> = cols.iter().map(|c| &c[..]).collect();
        let (status, chal, accum, folded) = bv.verify_with_hooks(&pr, &ins, &rlc).unwrap();
        let c = bv.info[3] as usize;
        for (j, p) in g["proofs"].as_array().unwrap().iter().enumerate() {
            assert_eq!(status[j] as u64, p["status"].as_u64().unwrap(), "{path:?} proof {j}: CUDA status");
            if let Some(acc) = p.get("accum") {
                let want: Vec<u8> = p["challenges"].as_array().unwrap().iter().flat_map(|v| fr_to_le(&fr_from_hex(v.as_str().unwrap()))).collect();
                assert_eq!(&chal[32 * c * j..32 * c * (j + 1)], &want[..], "{path:?} proof {j}: CUDA challenges");
                assert_eq!(hex::encode(&accum[128 * j..128 * (j + 1)]), acc.as_str().unwrap(), "{path:?} proof {j}: CUDA accumulators");
            }
        }
        assert_eq!(hex::encode(folded), g["folded"].as_str().unwrap(), "{path:?}: folded (L, R)");
        checked += 1;
    }
    assert!(checked >= 5, "golden vectors not found under H2V_B200_DIR/tests/golden");
}
