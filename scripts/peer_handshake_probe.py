"""Raw latency of the device-side cross-GPU handshakes of the peer transport (csrc/peer.h): two ranks, one CTA each,
rounds of enter+leave / leave only.  Launch one process per rank with RANK, WORLD_SIZE, JOB_ID (hex) set."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuadmm_b200 as cu
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); job = bytes.fromhex(os.environ["JOB_ID"])
dev = rank % max(cu.device_count(), 1)
out = (C.c_double * 3)()
rc = cu.lib.cuadmm_debug_peer_handshake(rank, world, job, dev, 2000, out)
print("rank", rank, "device", dev, "rc", rc, "us per round: enter+leave %.2f  leave %.2f  leave(volatile spin) %.2f" % (out[0], out[1], out[2]), flush=True)
