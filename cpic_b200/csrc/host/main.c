/* cpic_b200 command line: same usage as the reference's `cpic` (src/cpic.c) */
int cpic_b200_main(int argc, char **argv);
int main(int argc, char **argv) { return cpic_b200_main(argc, argv); }
