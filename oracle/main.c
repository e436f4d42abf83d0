/*
 * oracle/main.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 * Command dispatcher mirroring bamtk.c:227-319 for the three hot-path
 * commands, `bedcov` (a further caller of the same iterator, bedcov.c:303-331) plus `gl` (per-column genotype likelihoods).
 */
#include <stdio.h>
#include <string.h>
int main_mpileup(int argc, char **argv, int gl);
int main_depth(int argc, char **argv);
int main_coverage(int argc, char **argv);
int main_view(int argc, char **argv);
int main_bedcov(int argc, char **argv);
int main_pileup_dump(int argc, char **argv);
int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: plp_oracle mpileup|depth|coverage|gl [options]\n"); return 1; }
    if (!strcmp(argv[1], "mpileup")) return main_mpileup(argc - 1, argv + 1, 0);
    if (!strcmp(argv[1], "gl")) return main_mpileup(argc - 1, argv + 1, 1);
    if (!strcmp(argv[1], "depth")) return main_depth(argc - 1, argv + 1);
    if (!strcmp(argv[1], "coverage")) return main_coverage(argc - 1, argv + 1);
    if (!strcmp(argv[1], "bedcov")) return main_bedcov(argc - 1, argv + 1);
    if (!strcmp(argv[1], "view")) return main_view(argc - 1, argv + 1);
    if (!strcmp(argv[1], "pileup-dump")) return main_pileup_dump(argc - 1, argv + 1);
    fprintf(stderr, "plp_oracle: unknown command '%s'\n", argv[1]);
    return 1;
}
