//! Sigma1Handle (prove/src/sigma_source.rs:50-123) over device-resident tables: the four large components of sigma_1 are
//! uploaded once (from the mmapped rkyv archive, or a TZBWASM1 prover_crs section through tkm_crs_upload_mont) and every
//! encoder of the prover becomes one device MSM.
#![allow(non_snake_case)]
use crate::bivariate_polynomial::DensePolynomialExt;
use crate::group_structures::{G1Table, Sigma1Device};
use crate::iotools::{HexString, PlacementVariables, SetupParams, SubcircuitInfo};
use crate::{G1serde, ScalarField};

pub struct Sigma1Resident {
    pub xy_powers: Sigma1Device,
    pub gamma_inv_o_inst: G1Table,              // [l][1]
    pub eta_inv_li_o_inter_alpha4_kj: G1Table,  // [m_I][s_max]
    pub delta_inv_li_o_prv: G1Table,            // [m_D - l_D][s_max]
    pub delta: G1serde,
    pub eta: G1serde,
    pub delta_inv_alphak_xh_tx: Vec<Vec<G1serde>>,
    pub delta_inv_alpha4_xj_tx: Vec<G1serde>,
    pub delta_inv_alphak_yi_ty: Vec<Vec<G1serde>>,
}

pub struct Sigma1Handle<'a>(pub &'a Sigma1Resident);

impl<'a> Sigma1Handle<'a> {
    pub fn encode_poly(&self, poly: &mut DensePolynomialExt, _params: &SetupParams) -> G1serde { self.0.xy_powers.encode_poly(poly) }
    pub fn encode_poly_timed(&self, poly: &mut DensePolynomialExt, params: &SetupParams, _timing_name: &'static str) -> G1serde {
        self.encode_poly(poly, params)
    }

    /// encode_o_pub_free_common (group_structures/mod.rs:184-229): outputs of bufferPubOut, inputs of bufferPubIn / bufferBlockIn
    pub fn encode_O_pub_free(&self, placement_variables: &[PlacementVariables], subcircuit_infos: &[SubcircuitInfo], _p: &SetupParams) -> G1serde {
        let (mut idx, mut sc) = (Vec::new(), Vec::new());
        for pl in placement_variables {
            let info = &subcircuit_infos[pl.subcircuitId];
            let (start, cnt) = match info.name.as_str() {
                "bufferPubOut" => (info.Out_idx[0], info.Out_idx[1]),
                "bufferPubIn" | "bufferBlockIn" => (info.In_idx[0], info.In_idx[1]),
                _ => continue,
            };
            for j in start..start + cnt {
                idx.push(info.flattenMap[j] as u32);
                sc.push(pl.variables[j].to_scalar());
            }
        }
        self.0.gamma_inv_o_inst.msm_indexed(&sc, &idx)
    }

    /// encode_o_pub_fix_common (:145-182): the function instance against the last m_function entries of gamma_inv_o_inst
    pub fn encode_O_pub_fix(&self, a_pub_function: &[HexString], p: &SetupParams) -> G1serde {
        let m_function = p.l - p.l_free;
        if m_function == 0 { return G1serde::zero(); }
        if a_pub_function.len() != m_function {
            panic!("a_pub_function length mismatch: expected m_function={}, got a_pub_function.len()={}", m_function, a_pub_function.len());
        }
        let sc: Vec<ScalarField> = a_pub_function.iter().map(|h| h.to_scalar()).collect();
        let idx: Vec<u32> = (p.l - m_function..p.l).map(|j| j as u32).collect();
        self.0.gamma_inv_o_inst.msm_indexed(&sc, &idx)
    }

    /// encode_statement_common (:266-300) over the interface wires [l, l_D)
    pub fn encode_O_mid_no_zk(&self, pv: &[PlacementVariables], infos: &[SubcircuitInfo], p: &SetupParams) -> G1serde {
        Self::statement(&self.0.eta_inv_li_o_inter_alpha4_kj, pv, infos, p.l, p.l_D, p.s_max)
    }
    /// ... and over the private wires [l_D, m_D)
    pub fn encode_O_prv_no_zk(&self, pv: &[PlacementVariables], infos: &[SubcircuitInfo], p: &SetupParams) -> G1serde {
        Self::statement(&self.0.delta_inv_li_o_prv, pv, infos, p.l_D, p.m_D, p.s_max)
    }
    fn statement(table: &G1Table, pv: &[PlacementVariables], infos: &[SubcircuitInfo], lo: usize, hi: usize, s_max: usize) -> G1serde {
        let (mut idx, mut sc) = (Vec::new(), Vec::new());
        for (col, pl) in pv.iter().enumerate() {
            let info = &infos[pl.subcircuitId];
            for (local, &g) in info.flattenMap.iter().enumerate() {
                if g >= lo && g < hi {
                    idx.push(((g - lo) * s_max + col) as u32);
                    sc.push(pl.variables[local].to_scalar());
                }
            }
        }
        table.msm_indexed(&sc, &idx)
    }

    pub fn delta(&self) -> G1serde { self.0.delta }
    pub fn eta(&self) -> G1serde { self.0.eta }
    pub fn delta_inv_alphak_xh_tx(&self, k: usize, h: usize) -> G1serde { self.0.delta_inv_alphak_xh_tx[k][h] }
    pub fn delta_inv_alpha4_xj_tx(&self, j: usize) -> G1serde { self.0.delta_inv_alpha4_xj_tx[j] }
    pub fn delta_inv_alphak_yi_ty(&self, k: usize, i: usize) -> G1serde { self.0.delta_inv_alphak_yi_ty[k][i] }
}
