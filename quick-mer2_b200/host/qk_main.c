/*
 * qk_main.c -- command dispatch, the analogue of main() at Q.c:1496-1519 for the sub-commands
 * this build provides.  `quicKmer2_b200 count ...` and `quicKmer2_b200 est ...` take exactly the
 * arguments of `quicKmer2 count / est ...`; the offline dictionary builders (index, search, sparse)
 * stay with the reference binary.
 */
#include <stdio.h>
#include <string.h>

#include "../../include/qk_host.h"

int main(int argc, char **argv)
{
    if (argc >= 2 && strcmp(argv[1], "count") == 0) return qk_count_main(argc - 1, argv + 1); /* Q.c:1501 */
    if (argc >= 2 && strcmp(argv[1], "est") == 0) return qk_est_main(argc - 1, argv + 1);     /* Q.c:1507 */
    printf("%s\n\nquicKmer2_b200 count [-t N] [-g device] ref.fa sample.fast[a/q] Out_prefix\n"
           "quicKmer2_b200 est ref.fa sample_prefix output.bed\n"
           "(index, search and sparse are provided by the reference quicKmer2)\n",
           qk_version());
    return 1;
}
