//! Safe wrapper over `include/drprg_cuda.h`.
//!
//! Replaces `Pandora::genotype_with` (drprg `src/lib.rs:580-642`) for the map step of
//! `drprg predict` (`src/predict.rs:285-303`): same inputs (PRG, `--vcf-refs`, reads, outdir, the
//! `-t -w -k -c [-I] [-K]` options), same outputs (`<outdir>/pandora_genotyped.vcf`,
//! `<outdir>/pandora.log`), same error convention (an `Err` where the subprocess would have exited
//! non-zero).  NOTE: this crate is source only in this repository — the build container has no
//! Rust toolchain — and is kept deliberately small.
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_int};
use std::path::Path;

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct MapOpts {
    pub threads: u32,
    pub min_cluster_size: u32,
    pub illumina: u8,
    pub debug: u8,
    pub genome_size: u32,
    pub max_covg: u32,
    pub gt_conf: f64,
    pub genotyping_error_rate: f64,
    pub max_diff: u32,
    pub error_rate: f64,
}

#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct MapStats {
    pub n_reads: u64,
    pub n_reads_dropped: u64,
    pub total_bases: u64,
    pub n_hits: u64,
    pub n_hits_kept: u64,
    pub n_loci_present: u32,
    pub n_records: u32,
    pub exp_depth_covg: u32,
    pub ms_ingest: f64,
    pub ms_map: f64,
    pub ms_genotype: f64,
    pub ms_total: f64,
}

#[repr(C)]
pub struct RawIndex {
    _private: [u8; 0],
}

extern "C" {
    fn drprg_cuda_version() -> c_int;
    fn drprg_cuda_last_error() -> *const c_char;
    fn drprg_cuda_device_count() -> c_int;
    fn drprg_cuda_index_load(prg_path: *const c_char, w: u32, k: u32, device: c_int, out: *mut *mut RawIndex) -> c_int;
    fn drprg_cuda_index_free(idx: *mut RawIndex);
    fn drprg_cuda_map_genotype(
        idx: *mut RawIndex,
        reads_path: *const c_char,
        vcf_refs_fasta: *const c_char,
        outdir: *const c_char,
        opts: *const MapOpts,
        out_stats: *mut MapStats,
    ) -> c_int;
}

#[derive(thiserror::Error, Debug)]
pub enum CudaError {
    /// mirrors DependencyError::ProcessError (drprg src/lib.rs:65-90)
    #[error("drprg-cuda failed: {0}")]
    ProcessError(String),
    #[error("path is not valid UTF-8 / contains NUL: {0}")]
    BadPath(String),
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(drprg_cuda_last_error()).to_string_lossy().into_owned() }
}

fn cpath(p: &Path) -> Result<CString, CudaError> {
    CString::new(p.to_string_lossy().as_bytes()).map_err(|_| CudaError::BadPath(p.to_string_lossy().into_owned()))
}

pub fn version() -> i32 {
    unsafe { drprg_cuda_version() }
}

pub fn device_count() -> i32 {
    unsafe { drprg_cuda_device_count() }
}

/// PRG + k-mer graphs + minimizer table resident in HBM on one GPU.
pub struct Index {
    raw: *mut RawIndex,
}

unsafe impl Send for Index {}

impl Index {
    /// Parses and sketches the PRG (milliseconds) and uploads it; replaces the `dr.prg.kK.wW.idx` +
    /// `kmer_prgs/` files `pandora map` would reload.  Works for `outdir/updated.dr.prg` too.
    pub fn load(prg: &Path, w: u32, k: u32, device: i32) -> Result<Self, CudaError> {
        let p = cpath(prg)?;
        let mut raw: *mut RawIndex = std::ptr::null_mut();
        let rc = unsafe { drprg_cuda_index_load(p.as_ptr(), w, k, device, &mut raw) };
        if rc != 0 {
            return Err(CudaError::ProcessError(last_error()));
        }
        Ok(Index { raw })
    }

    /// Drop-in for `Pandora::genotype_with(prg, vcf_ref, reads, outdir, args)`.
    pub fn map_genotype(&mut self, vcf_ref: &Path, reads: &Path, outdir: &Path, opts: &MapOpts) -> Result<MapStats, CudaError> {
        let (r, v, o) = (cpath(reads)?, cpath(vcf_ref)?, cpath(outdir)?);
        let mut stats = MapStats::default();
        let rc = unsafe { drprg_cuda_map_genotype(self.raw, r.as_ptr(), v.as_ptr(), o.as_ptr(), opts, &mut stats) };
        if rc != 0 {
            return Err(CudaError::ProcessError(last_error()));
        }
        Ok(stats)
    }
}

impl Drop for Index {
    fn drop(&mut self) {
        unsafe { drprg_cuda_index_free(self.raw) }
    }
}

impl MapOpts {
    /// The options `Predict::run` passes to pandora (drprg src/predict.rs:288-294, src/lib.rs:594-609).
    pub fn for_predict(threads: u32, min_cluster_size: u32, illumina: bool, debug: bool) -> Self {
        MapOpts {
            threads,
            min_cluster_size,
            illumina: illumina as u8,
            debug: debug as u8,
            genome_size: 4_411_532, // MTB_GENOME_SIZE, src/lib.rs:36
            max_covg: u32::MAX,
            gt_conf: 0.0,
            genotyping_error_rate: 0.01,
            max_diff: 0,
            error_rate: 0.0,
        }
    }
}
