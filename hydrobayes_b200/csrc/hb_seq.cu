// hb_seq.cu -- sequential-sweep (dream(), R/dream.R:203-295) and batched independent-run mode: one CTA per run.
#include <string>

#include "hb_internal.h"

int hb_seq_supported(const hb_config* cfg, std::string& why) {
  (void)cfg;
  why = "sequential sweep / batched independent runs are not wired yet";
  return HB_ERR_UNSUPPORTED;
}
int hb_seq_run(hb_handle*, long long) { return HB_ERR_UNSUPPORTED; }
void hb_seq_destroy(hb_handle*) {}
