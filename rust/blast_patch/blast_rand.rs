//! Addition to blast/src/audio_processing/blast_rand.rs (the only change to that file).  X128P keeps generating on the
//! host for the REPL thread (commands.rs:838 seeds a generator per `seq` command); its state travels to the GPU inside
//! the Seq command, where every draw of the render is made (blast_command.rng_s0 / rng_s1).  The fields are private
//! (blast_rand.rs:4-8), so the patch adds one accessor inside `impl X128P`.  NOT compiled in the build image (no rustc).
impl X128P {
    /// (s0, s1) as SeqArgs.rng carries them to Conductor::apply
    pub fn state(&self) -> (u64, u64) { (self.s0, self.s1) }
}
