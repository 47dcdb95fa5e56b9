//! Plain-data views of the host-side JSON structures the commitment encoders read (libs/src/iotools/mod.rs:126-177,
//! 367-416): field names and meaning as in the reference; parsing stays host-only Rust (serde) and is not part of the
//! device path.  placementVariables.json can also be parsed in bulk by tkm_host_parse_hex_scalars.
#![allow(non_snake_case)]
use crate::ScalarField;

#[derive(Clone, Debug)]
pub struct HexString(pub String);
impl HexString {
    pub fn to_scalar(&self) -> ScalarField { ScalarField::from_hex(&self.0) }
}

#[derive(Clone, Debug)]
pub struct SetupParams {
    pub l_free: usize,
    pub l: usize,
    pub l_user_out: usize,
    pub l_user: usize,
    pub l_D: usize,
    pub m_D: usize,
    pub n: usize,
    pub s_D: usize,
    pub s_max: usize,
}

#[derive(Clone, Debug)]
pub struct SubcircuitInfo {
    pub id: usize,
    pub name: String,
    pub Nwires: usize,
    pub Nconsts: usize,
    pub Out_idx: [usize; 2],
    pub In_idx: [usize; 2],
    pub flattenMap: Vec<usize>,
}

#[derive(Clone, Debug)]
pub struct PlacementVariables {
    pub subcircuitId: usize,
    pub variables: Vec<HexString>,
}
