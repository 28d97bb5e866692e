"""Mint tests/golden/diffusers_keys.json from the reference's own converter script: the diffusers-side key templates are
read out of tools/convert_pixart_to_diffusers.py (the assignments `converted_state_dict[<key>] = ...` and the PixArt-side
`state_dict.pop(<key>)` names), so instarevive_b200/convert.py is pinned to the names the reference writes, not to a
restatement of them. Run:  python oracle/make_goldens_convert.py"""
import json
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/tools/convert_pixart_to_diffusers.py").read_text()

body = SRC[SRC.index("def main(args):"):SRC.index("# PixArt XL/2")]
dst = re.findall(r'converted_state_dict\[f?"([^"]+)"\]', body)
src = re.findall(r'state_dict\.pop\(\s*f?"([^"]+)"', body)
assert dst and src
out = {"diffusers_keys": sorted(set(dst)), "pixart_keys": sorted(set(src)),
       "source": "tools/convert_pixart_to_diffusers.py:29-154 ({depth} = block index 0..27)"}
(ROOT / "tests" / "golden" / "diffusers_keys.json").write_text(json.dumps(out, indent=1))
print(len(out["diffusers_keys"]), "diffusers key templates,", len(out["pixart_keys"]), "pixart key templates")
