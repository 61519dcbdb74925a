/* TEST INFRASTRUCTURE -- not part of the product.
 * Tiny front-end to the reference's vendored libbam (samtools-0.1.18), built by oracle/Makefile.ref:
 *   bamtool index  in.bam            -> in.bam.bai   (bam_index_build, bam_index.c)
 *   bamtool sort   in.bam out_prefix -> out_prefix.bam (bam_sort_core, bam_sort.c)
 * Used to index the synthetic BAMs our own writer (rsicnv_b200/synth.py) produces, so that the
 * reference CLI can read them; doubles as a cross-check of that writer against libbam's reader. */
#include <stdio.h>
#include <string.h>
#include "bam.h"
void bam_sort_core(int is_by_qname, const char *fn, const char *prefix, size_t max_mem);
int main(int argc, char **argv) {
  if (argc >= 3 && strcmp(argv[1], "index") == 0) return bam_index_build(argv[2]);
  if (argc >= 4 && strcmp(argv[1], "sort") == 0) { bam_sort_core(0, argv[2], argv[3], 500000000); return 0; }
  fprintf(stderr, "usage: bamtool index in.bam | bamtool sort in.bam out_prefix\n");
  return 2;
}
