"""Host-side protocol layer over the device-resident engine: the callers either side of the hot path.

  formats     file formats the reference reads/writes (setupParams/subcircuitInfo/placementVariables/permutation/
              instance JSON, iden3 .r1cs binaries, the Solidity-split proof/preprocess JSON)
  qap         sparse R1CS x witness products (u/v/w evaluation tables), o_j(tau) of the QAP mixture
  setup       Tau, Sigma1/Sigma2 generation (`trusted-setup --fixed-tau`; fixed-base multiples on the GPU)
  prover      Prover.init / prove0..prove4 + the Keccak transcript (prove/src/lib.rs)
  preprocess  s0, s1, O_pub_fix (preprocess/src/lib.rs)
  verifier    the pairing acceptance check (verify-rust/src/lib.rs)
  pairing     BLS12-381 optimal ate pairing on the host (CPU, like arkworks in the reference)
  synthetic   a generator of forward-solvable subcircuit libraries + synthesizer outputs of the reference's shapes

The driver is written once against a small backend interface (polynomials, commitments, sparse MSMs, G1 ops); the
product backend is `GpuBackend` (libtokamak_b200 through the C-ABI).  Tests plug the CPU oracle in through the same
interface to check byte-identical proofs."""
