//! Drop-in replacement of the hot path of rostin79s/TFHE-AES (`src/server/`): the same `Server` methods and the same
//! `sbox` module functions, executed by libtfhe_aes_b200.so on a B200 instead of tfhe-rs' CPU WoPBS.
//!
//! ```text
//! reference                                        this crate
//! src/server/server.rs     Server                  server::Server                     (same method signatures)
//! src/server/sbox/gen_lut.rs, sbox.rs, many_wopbs.rs   sbox::{gen_lut, sbox, many_sbox, many_wopbs_without_padding}
//! src/main.rs:55-64        rayon loop              Server::aes_ctr  (one batched call; the per-block calls also work)
//! ```
//! `src/client/client.rs` and `src/main.rs` stay as they are.
pub mod ffi;
pub mod flatten;
pub mod sbox;
pub mod server;

pub use server::Server;
