"""Golden digests for the PNG writer: sha256 of stbi_write_png_to_mem (the reference's vendored stb_image_write v1.16,
compiled where it lies by oracle/build_ref.sh) on the seeded images of tests/png_cases.py.  Run in the build container
(needs /root/reference); the digests travel to the GPU box, the reference does not."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
import png_cases  # noqa: E402

out = {name: {"shape": list(img.shape), "sha256": hashlib.sha256(O.ref_png(img)).hexdigest(), "bytes": len(O.ref_png(img))}
       for name, img in png_cases.images()}
json.dump(out, open(os.path.join(HERE, "png_stb.json"), "w"), indent=1)
print(f"wrote {len(out)} digests")
