#!/bin/bash
# round 2, GPU call 33: round-end check after the ChaCha20 generator of the vanishing argument's random polynomial
# (every proof's bytes change, on the oracle's side and on the device's: the whole GPU suite, smoke, the bench line)
cd "$(dirname "$0")/../.."
bash tools/gpu_round_check.sh
