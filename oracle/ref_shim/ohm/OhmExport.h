// Stand-in for the CMake-generated export header (generate_export_header): static build, default visibility.
#ifndef OHM_EXPORT_H
#define OHM_EXPORT_H
#define ohm_API
#define OHM_NO_EXPORT
#define ohm_DEPRECATED __attribute__((__deprecated__))
#endif
