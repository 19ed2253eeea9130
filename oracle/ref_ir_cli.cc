// oracle/ref_ir_cli.cc — TEST INFRASTRUCTURE.  Prints the reference's lowered
// IR (RegexpInfo lists, src/regexp.h:538-636) for one pattern by switching on
// its own --print_re_list / --print_re_tree flags (src/flags.h:24-34) and
// compiling for kMatchAll with fast-forward off.  Used by
// tests/golden/make_golden.py to pin the oracle's state numbering.
#include <stdio.h>
#include <stdlib.h>
#include "rejit.h"
#include "flags.h"
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  int parser_opt = argc > 2 ? atoi(argv[2]) : 1;
  FLAG_use_fast_forward = false;
  FLAG_use_fast_forward_early = false;
  FLAG_use_ff_reduce = false;  // FF_finder would otherwise append "linking" nodes to the lists (src/codegen.cc:395-531)
  FLAG_use_parser_opt = parser_opt != 0;
  FLAG_print_re_list = true;
  FLAG_print_state_ring_info = true;
  rejit::Regej re(argv[1]);
  if (re.status() != rejit::RejitSuccess) { printf("PARSE_ERROR\n"); return 1; }
  re.Compile(rejit::kMatchAll);
  return 0;
}
