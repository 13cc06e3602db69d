/*
 * qk_main.c -- command dispatch, the analogue of main() at Q.c:1496-1519 for the one
 * sub-command this build provides.  `quicKmer2_b200 count ...` takes exactly the
 * arguments of `quicKmer2 count ...`; the other sub-commands (index, search, est, sparse)
 * stay with the reference binary.
 */
#include <stdio.h>
#include <string.h>

#include "../../include/qk_host.h"

int main(int argc, char **argv)
{
    if (argc >= 2 && strcmp(argv[1], "count") == 0) return qk_count_main(argc - 1, argv + 1); /* Q.c:1501 */
    printf("%s\n\nquicKmer2_b200 count [-t N] [-g device] ref.fa sample.fast[a/q] Out_prefix\n"
           "(index, search, est and sparse are provided by the reference quicKmer2)\n",
           qk_version());
    return 1;
}
