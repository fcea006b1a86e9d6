#!/usr/bin/env bash
# Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, copied byte for byte from
# /root/reference (read-only mount of wazzuck/multimodal-image-transformer) into the git-ignored oracle/_ref/.
#
#   decoder.py   TransformerDecoder / PositionalEncodingBatchFirst   (the hot path, decoder.py:16-193)
#   utils.py     generate_square_subsequent_mask / create_padding_mask (utils.py:11-70)
#   config.py    module constants both of them import               (config.py:1-145)
#
# oracle/_ref/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs
# (`cpu_baseline`, `--impl reference`) import it -- as the checker / the reported baseline, never as the product.
# It is listed in .gitignore (sources stay out of history) and NOT in .gpurunignore, so it travels to the GPU box
# next to the built .so; `/root/reference` itself does not exist there.  __graft_entry__.build() runs this script
# whenever /root/reference is mounted.  A sha256 manifest is written so that a stale or edited copy is detected.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "make_ref.sh: $REF not mounted; keeping whatever is in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
for f in decoder.py utils.py config.py; do
  cp -f "$REF/$f" "$OUT/$f" && chmod u+w "$OUT/$f"
done
( cd "$OUT" && sha256sum decoder.py utils.py config.py > MANIFEST.sha256 )
( cd "$REF" && sha256sum decoder.py utils.py config.py ) | diff -q - "$OUT/MANIFEST.sha256" >/dev/null
echo "oracle/_ref: $(wc -l < "$OUT/decoder.py") + $(wc -l < "$OUT/utils.py") + $(wc -l < "$OUT/config.py") lines copied unmodified from $REF"
