// Stand-in for the CMake-generated export header of ohmgpu (generate_export_header): default visibility.
#ifndef OHMGPU_EXPORT_H
#define OHMGPU_EXPORT_H
#define ohmgpu_API
#define OHMGPU_NO_EXPORT
#define ohmgpu_DEPRECATED __attribute__((__deprecated__))
#endif
