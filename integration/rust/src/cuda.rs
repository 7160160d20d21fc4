//! FFI declarations (exactly the symbols of include/h2v.h that a host needs) and the safe wrapper.
use core::ffi::c_char;
use ff::PrimeField;
use halo2_verifier::{
    halo2curves::bn256::{Fr, G1Affine},
    helpers::SerdeFormat,
    plonk::Error,
    ParamsKZG, VerifyingKey,
};

#[repr(C)]
pub struct H2vCtx {
    _p: [u8; 0],
}

pub const H2V_COMM_HANDLE_BYTES: usize = 128;

extern "C" {
    fn h2v_ctx_create_multi(out: *mut *mut H2vCtx, params: *const u8, params_len: usize, params_format: i32, vk: *const u8, vk_len: usize,
                            vk_format: i32, multiopen: i32, hash: i32, device: i32, circuit_instances: u32) -> i32;
    fn h2v_ctx_destroy(ctx: *mut H2vCtx);
    fn h2v_last_error(ctx: *const H2vCtx) -> *const c_char;
    fn h2v_verify_batch(ctx: *mut H2vCtx, n: u32, proofs: *const u8, proof_off: *const u64, instances: *const u8, inst_off: *const u64,
                        rlc_scalars: *const u8, seed: u64, status: *mut u8, challenges: *mut u8, accum: *mut u8, batch_accum: *mut u8) -> i32;
    fn h2v_batch_set_columns(ctx: *mut H2vCtx, inst_ncols: *const u32, inst_col_len: *const u32) -> i32;
    fn h2v_batch_set_fold_groups(ctx: *mut H2vCtx, groups: u32) -> i32;
    fn h2v_last_group_verdicts(ctx: *const H2vCtx, out: *mut u8, capacity: u32) -> i32;
    fn h2v_ctx_info(ctx: *const H2vCtx, out8: *mut u32) -> i32;
    // sharded batches over the GPUs of one node (device-side exchange, include/h2v.h)
    fn h2v_comm_init(ctx: *mut H2vCtx, rank: u32, world: u32, max_groups: u32, handle_out: *mut u8) -> i32;
    fn h2v_comm_connect(ctx: *mut H2vCtx, handles: *const u8) -> i32;
    fn h2v_batch_set_rlc_key(ctx: *mut H2vCtx, key32: *const u8) -> i32;
    fn h2v_verify_shard(ctx: *mut H2vCtx, n: u32, proofs: *const u8, proof_off: *const u64, instances: *const u8, inst_off: *const u64,
                        rlc_scalars: *const u8, seed: u64, global_base: u64, global_count: u64, root: u32, status: *mut u8,
                        group_verdicts: *mut u8, verdict: *mut i32) -> i32;
}

/// generic parameter `V` of the reference's `verify_proof` (lib.rs:36): `VerifierSHPLONK` / `VerifierGWC`
#[derive(Clone, Copy)]
pub enum MultiOpen {
    Shplonk = 0,
    Gwc = 1,
}
/// generic parameter `T` (lib.rs:38): `Blake2bRead` / `Keccak256Read`, both with `Challenge255`
#[derive(Clone, Copy)]
pub enum TranscriptHash {
    Blake2b = 0,
    Keccak256 = 1,
}

/// One (params, vk, scheme, transcript, device) context: owns the device plan, a stream and all batch buffers.
/// Single owner, one batch in flight; use one per host thread.
pub struct BatchVerifier {
    ctx: *mut H2vCtx,
    circuit_instances: usize,
    /// out8 of h2v_ctx_info: k, points, scalars, challenges, proof length, instance columns, shared bases, multi-open points
    pub info: [u32; 8],
}
unsafe impl Send for BatchVerifier {}

/// status byte -> what `verify_proof` returns for that proof (plonk/mod.rs:19-32)
pub fn status_to_result(s: u8) -> Result<(), Error> {
    match s {
        0 => Ok(()),
        1 => Err(Error::InvalidInstances),
        2 => Err(Error::Transcript(halo2_verifier::io::Error::new(halo2_verifier::io::ErrorKind::Other, "transcript read failed"))),
        3 => Err(Error::Opening),
        4 => Err(Error::ConstraintSystemFailure),
        _ => panic!("input on which the reference verifier panics (vanishing.rs:100 / shplonk.rs:215)"),
    }
}

struct Packed {
    pbuf: Vec<u8>,
    poff: Vec<u64>,
    ibuf: Vec<u8>,
    ioff: Vec<u64>,
    ncols: Vec<u32>,
    col_len: Vec<u32>,
    ragged: bool,
}

impl BatchVerifier {
    pub fn new(params: &ParamsKZG, vk: &VerifyingKey<G1Affine>, mo: MultiOpen, th: TranscriptHash, device: i32) -> Result<Self, String> {
        Self::new_multi(params, vk, mo, th, device, 1)
    }

    /// `circuit_instances` = `instances.len()` of one `verify_proof` call (lib.rs:63,92,117,134): proofs that carry
    /// several circuit instances in one transcript
    pub fn new_multi(params: &ParamsKZG, vk: &VerifyingKey<G1Affine>, mo: MultiOpen, th: TranscriptHash, device: i32, circuit_instances: usize)
        -> Result<Self, String> {
        let (mut pb, mut vb) = (Vec::new(), Vec::new());
        params.write_custom(&mut pb, SerdeFormat::RawBytes).map_err(|e| format!("{e:?}"))?; // poly/kzg/commitment.rs:142-152
        vk.write(&mut vb, SerdeFormat::RawBytes).map_err(|e| format!("{e:?}"))?; // plonk/vk.rs:41-64
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe {
            h2v_ctx_create_multi(&mut ctx, pb.as_ptr(), pb.len(), 1, vb.as_ptr(), vb.len(), 1, mo as i32, th as i32, device, circuit_instances as u32)
        };
        if rc != 0 {
            return Err(last_error(core::ptr::null()));
        }
        let mut info = [0u32; 8];
        unsafe { h2v_ctx_info(ctx, info.as_mut_ptr()) };
        Ok(Self { ctx, circuit_instances, info })
    }

    /// `instances[proof][circuit instance][column][row]` for multi-instance contexts is flattened by the caller to
    /// `instances[proof][instance-major column][row]`; for the usual single-instance case it is `instances[proof][column][row]`.
    fn pack(&self, proofs: &[&[u8]], instances: &[&[&[Fr]]]) -> Packed {
        let (mut pbuf, mut poff) = (Vec::new(), vec![0u64]);
        for p in proofs {
            pbuf.extend_from_slice(p);
            poff.push(pbuf.len() as u64);
        }
        let cols = self.info[5] as usize;
        let (mut ibuf, mut ioff, mut ncols, mut col_len, mut ragged) = (Vec::new(), vec![0u64], Vec::new(), Vec::new(), false);
        for inst in instances {
            ncols.push(inst.len() as u32);
            ragged |= inst.len() != cols || inst.iter().any(|c| c.len() != inst[0].len());
            for c in 0..cols {
                let col: &[Fr] = inst.get(c).copied().unwrap_or(&[]);
                col_len.push(col.len() as u32);
            }
            for col in inst.iter() {
                for v in col.iter() {
                    ibuf.extend_from_slice(v.to_repr().as_ref()); // 32-byte little-endian canonical, as transcript/mod.rs:228
                }
            }
            ioff.push((ibuf.len() / 32) as u64);
        }
        Packed { pbuf, poff, ibuf, ioff, ncols, col_len, ragged }
    }

    /// One entry per proof: `Ok(())` or the error the reference's `verify_proof` returns.  The fold coefficients are
    /// drawn from the OS inside the library (seed 0, no scalars), as the reference does (strategy.rs:129).
    pub fn verify_proofs_batch(&mut self, proofs: &[&[u8]], instances: &[&[&[Fr]]]) -> Result<Vec<Result<(), Error>>, String> {
        assert_eq!(proofs.len(), instances.len());
        let _ = self.circuit_instances;
        let p = self.pack(proofs, instances);
        let n = proofs.len();
        let mut status = vec![0u8; n];
        unsafe {
            if p.ragged {
                h2v_batch_set_columns(self.ctx, p.ncols.as_ptr(), if self.info[5] > 0 { p.col_len.as_ptr() } else { core::ptr::null() });
            }
            let rc = h2v_verify_batch(self.ctx, n as u32, p.pbuf.as_ptr(), p.poff.as_ptr(), p.ibuf.as_ptr(), p.ioff.as_ptr(), core::ptr::null(), 0,
                                      status.as_mut_ptr(), core::ptr::null_mut(), core::ptr::null_mut(), core::ptr::null_mut());
            if rc != 0 {
                return Err(last_error(self.ctx));
            }
        }
        Ok(status.into_iter().map(status_to_result).collect())
    }

    /// Parity hooks (differential test): explicit fold scalars r_i in, per-proof statuses, transcript challenges
    /// (n x C x 32 B), per-proof affine accumulators (n x 128 B: L_j | R_j) and the folded (L | R) out.
    pub fn verify_with_hooks(&mut self, proofs: &[&[u8]], instances: &[&[&[Fr]]], rlc: &[Fr]) -> Result<(Vec<u8>, Vec<u8>, Vec<u8>, [u8; 128]), String> {
        let p = self.pack(proofs, instances);
        let n = proofs.len();
        let c = self.info[3] as usize;
        let rbytes: Vec<u8> = rlc.iter().flat_map(|r| r.to_repr().as_ref().to_vec()).collect();
        let (mut status, mut chal, mut accum, mut folded) = (vec![0u8; n], vec![0u8; 32 * n * c], vec![0u8; 128 * n], [0u8; 128]);
        unsafe {
            if p.ragged {
                h2v_batch_set_columns(self.ctx, p.ncols.as_ptr(), if self.info[5] > 0 { p.col_len.as_ptr() } else { core::ptr::null() });
            }
            let rc = h2v_verify_batch(self.ctx, n as u32, p.pbuf.as_ptr(), p.poff.as_ptr(), p.ibuf.as_ptr(), p.ioff.as_ptr(), rbytes.as_ptr(), 0,
                                      status.as_mut_ptr(), chal.as_mut_ptr(), accum.as_mut_ptr(), folded.as_mut_ptr());
            if rc != 0 {
                return Err(last_error(self.ctx));
            }
        }
        Ok((status, chal, accum, folded))
    }

    /// Throughput mode: `groups` consecutive independent batches (own fold, own pairing check, own verdict) through one
    /// set of kernel launches; returns the per-proof results and the per-batch verdicts.
    pub fn verify_batches(&mut self, groups: u32, proofs: &[&[u8]], instances: &[&[&[Fr]]]) -> Result<(Vec<Result<(), Error>>, Vec<bool>), String> {
        unsafe { h2v_batch_set_fold_groups(self.ctx, groups) };
        let res = self.verify_proofs_batch(proofs, instances)?;
        let mut gv = vec![0u8; groups as usize];
        unsafe { h2v_last_group_verdicts(self.ctx, gv.as_mut_ptr(), groups) };
        Ok((res, gv.into_iter().map(|v| v != 0).collect()))
    }
}

impl Drop for BatchVerifier {
    fn drop(&mut self) {
        unsafe { h2v_ctx_destroy(self.ctx) }
    }
}

/// One process per GPU: the device-side exchange channel of a context (include/h2v.h, "device-side exchange").
/// `export` -> ship the 128-byte handle of every rank to every rank (any transport) -> `connect`.
pub struct ShardChannel<'a> {
    bv: &'a mut BatchVerifier,
    pub rank: u32,
    pub world: u32,
}

impl<'a> ShardChannel<'a> {
    pub fn export(bv: &'a mut BatchVerifier, rank: u32, world: u32, max_groups: u32) -> Result<(Self, [u8; H2V_COMM_HANDLE_BYTES]), String> {
        let mut h = [0u8; H2V_COMM_HANDLE_BYTES];
        if unsafe { h2v_comm_init(bv.ctx, rank, world, max_groups, h.as_mut_ptr()) } != 0 {
            return Err(last_error(bv.ctx));
        }
        Ok((Self { bv, rank, world }, h))
    }
    /// handles of all ranks in rank order (own included)
    pub fn connect(&mut self, handles: &[[u8; H2V_COMM_HANDLE_BYTES]]) -> Result<(), String> {
        let flat: Vec<u8> = handles.iter().flat_map(|h| h.to_vec()).collect();
        if unsafe { h2v_comm_connect(self.bv.ctx, flat.as_ptr()) } != 0 {
            return Err(last_error(self.bv.ctx));
        }
        Ok(())
    }
    /// This rank's shard [global_base, global_base + proofs.len()) of a global batch of `global_count` proofs; `key` =
    /// one fresh 32-byte secret per global batch, the same on every rank (the fold coefficients are then defined
    /// globally); `root` = the rank that sums the partial accumulators and runs the single pairing check.
    pub fn verify_shard(&mut self, proofs: &[&[u8]], instances: &[&[&[Fr]]], global_base: u64, global_count: u64, root: u32, key: &[u8; 32])
        -> Result<(bool, Vec<Result<(), Error>>), String> {
        let p = self.bv.pack(proofs, instances);
        let n = proofs.len();
        let (mut status, mut verdict) = (vec![0u8; n], 0i32);
        unsafe {
            h2v_batch_set_rlc_key(self.bv.ctx, key.as_ptr());
            let rc = h2v_verify_shard(self.bv.ctx, n as u32, p.pbuf.as_ptr(), p.poff.as_ptr(), p.ibuf.as_ptr(), p.ioff.as_ptr(), core::ptr::null(), 0,
                                      global_base, global_count, root, status.as_mut_ptr(), core::ptr::null_mut(), &mut verdict);
            if rc != 0 {
                return Err(last_error(self.bv.ctx));
            }
        }
        Ok((verdict == 1, status.into_iter().map(status_to_result).collect()))
    }
}

fn last_error(ctx: *const H2vCtx) -> String {
    unsafe { std::ffi::CStr::from_ptr(h2v_last_error(ctx)) }.to_string_lossy().into_owned()
}
