// FrayTracer.B200.fs -- P/Invoke shim over libfraytracer_b200.so (include/fraytracer_b200.h).
//
// UNVERIFIED SOURCE: no .NET toolchain exists in the build environment of this repository, so this file has
// never been compiled.  It shows the binding a FrayTracer maintainer would add: the same constructor names as
// SdfForm / SdfObject / SdfMaterial / SdfLight (src/FrayTracer/*.fs), returning reified nodes instead of
// closures, and drop-in replacements for `SdfScene.trace` + `Image.render` (Program.fs:90-93).
namespace FrayTracer.B200

open System
open System.Numerics
open System.Runtime.InteropServices
open FrayTracer

#nowarn "9"

module Native =
    [<Literal>]
    let Lib = "fraytracer_b200"

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type FormDesc =
        val mutable kind : int
        val mutable firstChild : int
        val mutable childCount : int
        [<MarshalAs(UnmanagedType.ByValArray, SizeConst = 10)>]
        val mutable p : float32[]

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type MaterialDesc = { kind : int; r : float32; g : float32; b : float32 }

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type ObjectDesc = { kind : int; form : int; material : int; firstChild : int; childCount : int }

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type LightDesc = { kind : int; vx : float32; vy : float32; vz : float32; r : float32; g : float32; b : float32 }

    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type SceneDesc =
        { forms : nativeint; objects : nativeint; children : nativeint; materials : nativeint; lights : nativeint
          nForms : int; nObjects : int; nChildren : int; nMaterials : int; nLights : int; rootObject : int
          bgR : float32; bgG : float32; bgB : float32 }

    // ft_cloud_params (EXTENSION): perm -> 256 x int32, gradients -> 256 x 3 x float (F# float = float64)
    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type CloudParamsDesc =
        { density : float32; frequency : float32; step : float32; fade : float32; bias : float32; gain : float32
          perm : nativeint; gradients : nativeint }

    // mirrors the Camera record (Camera.fs:16-22): 4 x Vector3
    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type CameraDesc = { position : Vector3; forward : Vector3; upScaled : Vector3; rightScaled : Vector3 }

    [<DllImport(Lib)>] extern int ft_abi_version()
    [<DllImport(Lib)>] extern nativeint ft_last_error()
    [<DllImport(Lib)>] extern int ft_scene_create(SceneDesc& desc, int device, nativeint& scene)
    [<DllImport(Lib)>] extern int ft_scene_destroy(nativeint scene)
    [<DllImport(Lib)>] extern int ft_host_register(nativeint ptr, int64 bytes)
    [<DllImport(Lib)>] extern int ft_host_unregister(nativeint ptr)
    [<DllImport(Lib)>]
    extern int ft_render(nativeint scene, CameraDesc& camera, int width, int height, float32 epsilon, float32 length,
                         uint32 flags, nativeint outRgb, nativeint outLeaf, nativeint stats)
    // EXTENSIONS (no reference implementation; DESIGN.md 4.6 / 4.7)
    [<DllImport(Lib)>]
    extern int ft_render_spectral(nativeint scene, CameraDesc& camera, int width, int height, float32 epsilon,
                                  float32 length, int nWavelengths, int maxBounces, uint64 seed, nativeint outRgb,
                                  nativeint stats)
    [<DllImport(Lib)>]
    extern int ft_render_cloud(nativeint scene, CameraDesc& camera, int width, int height, float32 epsilon,
                               float32 length, CloudParamsDesc& cloud, int samplesPerPixel, uint64 seed,
                               nativeint outRgb, nativeint stats)
    [<DllImport(Lib)>]
    extern int ft_trace_rays(nativeint scene, int64 n, nativeint rays, uint32 flags, nativeint outRgb,
                             nativeint outLeaf, nativeint outSteps, nativeint outShadow, nativeint stats)

    let check (status:int) =
        if status <> 0 then failwith (Marshal.PtrToStringAnsi(ft_last_error()))

// ---- reified nodes: same constructors as the reference, data instead of closures ----------------------------
type FormNode =
    | Sphere of SdfForm.Primitive.Sphere          // SdfForm.fs:118-135
    | Capsule of SdfForm.Primitive.Capsule        // SdfForm.fs:137-170
    | Torus of SdfForm.Primitive.Torus            // SdfForm.fs:172-203
    | Triangle of SdfForm.Primitive.Triangle      // SdfForm.fs:205-268
    | Union of FormNode[]                         // SdfForm.fs:14-40
    | Subtract of FormNode * FormNode             // SdfForm.fs:42-49
    | Intersect of FormNode[]                     // SdfForm.fs:51-67
    | UnionSmooth of float32 * FormNode[]         // SdfForm.fs:69-91

type MaterialNode = Solid of FColor              // SdfMaterial.fs:4-7

type ObjectNode =
    | Leaf of MaterialNode * FormNode             // SdfObject.create   SdfObject.fs:6-10
    | ObjUnion of ObjectNode[]                    // SdfObject.union    SdfObject.fs:12-48
    | ObjSubtract of ObjectNode * FormNode        // SdfObject.subtract SdfObject.fs:50-54
    | ObjIntersect of ObjectNode * FormNode[]     // SdfObject.intersect SdfObject.fs:56-64

type LightNode =
    | Directional of Vector3 * FColor             // SdfLight.directional SdfLight.fs:6-21
    | Point of Vector3 * FColor                   // SdfLight.point       SdfLight.fs:23-42

type SceneNode = { Object : ObjectNode; BackgroundColor : FColor; Lights : LightNode list }   // Types.fs:74-79

[<RequireQualifiedAccess>]
module SdfForm =
    module Primitive =
        let sphere d = Sphere d
        let capsule d = Capsule d
        let torus d = Torus d
        let triangle d = Triangle d
    let union (forms:seq<FormNode>) =
        match Seq.toArray forms with
        | [||] -> failwith "No SdfObjects given."
        | [|f|] -> f
        | fs -> Union fs
    let subtract a b = Subtract (a, b)
    let intersect (forms:seq<FormNode>) =
        match Seq.toArray forms with
        | [||] -> failwith "No SdfObjects given."
        | [|f|] -> f
        | fs -> Intersect fs
    let unionSmooth strength (forms:seq<FormNode>) =
        match Seq.toArray forms with
        | [||] -> failwithf "blub"
        | [|f|] -> f
        | fs -> UnionSmooth (strength, fs)

[<RequireQualifiedAccess>]
module SdfMaterial =
    let createSolid color = Solid color

[<RequireQualifiedAccess>]
module SdfObject =
    let create material form = Leaf (material, form)
    let union (objects:seq<ObjectNode>) =
        match Seq.toArray objects with
        | [||] -> failwith "No SdfObjects given."
        | [|o|] -> o
        | os -> ObjUnion os
    let subtract o form = ObjSubtract (o, form)
    let intersect o (forms:seq<FormNode>) = ObjIntersect (o, Seq.toArray forms)

[<RequireQualifiedAccess>]
module SdfLight =
    let directional direction color = Directional (direction, color)
    let point position color = Point (position, color)

/// A scene compiled onto one B200.  `SdfScene.trace scene` of the reference becomes `new CompiledScene(scene, 0)`.
/// The node graph is flattened into the tables of ft_scene_desc (children before parents, one shared `children`
/// index array) -- the same walk as fraytracer_b200/api.py : SceneDescription.__init__, the tested binding.
type CompiledScene(scene:SceneNode, device:int) =
    let mutable handle = 0n
    do
        let forms = ResizeArray<int * int * int * float32[]>()        // kind, firstChild, childCount, p[10]
        let objects = ResizeArray<int * int * int * int * int>()      // kind, form, material, firstChild, childCount
        let children = ResizeArray<int>()
        let materials = ResizeArray<FColor>()
        let v (a:Vector3) = [ a.X; a.Y; a.Z ]
        let pad (xs:float32 list) = Array.append (Array.ofList xs) (Array.zeroCreate (10 - xs.Length))
        let rec addForm (f:FormNode) : int =
            let prim kind p = forms.Add((kind, 0, 0, pad p)); forms.Count - 1
            let composite kind (kids:FormNode[]) p =
                let ids = kids |> Array.map addForm
                let first = children.Count
                children.AddRange ids
                forms.Add((kind, first, ids.Length, pad p))
                forms.Count - 1
            match f with
            | Sphere s -> prim 1 (v s.Center @ [ s.Radius ])                                         // FT_FORM_SPHERE
            | Capsule c -> prim 2 (v c.From @ v c.To @ [ c.Radius ])                                 // FT_FORM_CAPSULE
            | Torus t -> prim 3 (v t.Center @ v t.Normal @ [ t.MajorRadius; t.MinorRadius ])         // FT_FORM_TORUS
            | Triangle t -> prim 4 (v t.V1 @ v t.V2 @ v t.V3 @ [ t.Radius ])                         // FT_FORM_TRIANGLE
            | Union fs -> composite 16 fs []
            | Subtract (a, b) -> composite 17 [| a; b |] []
            | Intersect fs -> composite 18 fs []
            | UnionSmooth (k, fs) -> composite 19 fs [ k ]
        let rec addObject (o:ObjectNode) : int =
            match o with
            | Leaf (Solid color, form) ->
                let fid = addForm form
                materials.Add color
                objects.Add((1, fid, materials.Count - 1, 0, 0))                                     // FT_OBJECT_LEAF
            | ObjUnion os ->
                let ids = os |> Array.map addObject
                let first = children.Count
                children.AddRange ids
                objects.Add((2, -1, -1, first, ids.Length))                                          // FT_OBJECT_UNION
            | ObjSubtract (b, form) ->
                let bid = addObject b
                let fid = addForm form
                let first = children.Count
                children.Add bid
                objects.Add((3, fid, -1, first, 1))                                                  // FT_OBJECT_SUBTRACT
            | ObjIntersect (b, fs) ->
                let bid = addObject b
                let fids = fs |> Array.map addForm
                let first = children.Count
                children.Add bid
                children.AddRange fids
                objects.Add((4, -1, -1, first, 1 + fids.Length))                                     // FT_OBJECT_INTERSECT
            objects.Count - 1
        let root = addObject scene.Object
        // unmanaged copies of the tables (ft_form_desc is 52 bytes, ft_object_desc 20, ft_material_desc 16, ft_light_desc 28)
        let alloc n = Marshal.AllocHGlobal(max 1 n)
        let pForms, pObjects = alloc (52 * forms.Count), alloc (20 * objects.Count)
        let pChildren, pMaterials = alloc (4 * children.Count), alloc (16 * materials.Count)
        let lights = List.toArray scene.Lights
        let pLights = alloc (28 * lights.Length)
        let f32 (p:nativeint) (off:int) (x:float32) = Marshal.WriteInt32(p, off, BitConverter.SingleToInt32Bits x)
        try
            forms |> Seq.iteri (fun i (kind, first, count, p) ->
                Marshal.WriteInt32(pForms, 52 * i, kind)
                Marshal.WriteInt32(pForms, 52 * i + 4, first)
                Marshal.WriteInt32(pForms, 52 * i + 8, count)
                Marshal.Copy(p, 0, pForms + nativeint (52 * i + 12), 10))
            objects |> Seq.iteri (fun i (kind, form, material, first, count) ->
                [ kind; form; material; first; count ] |> List.iteri (fun k x -> Marshal.WriteInt32(pObjects, 20 * i + 4 * k, x)))
            children |> Seq.iteri (fun i c -> Marshal.WriteInt32(pChildren, 4 * i, c))
            materials |> Seq.iteri (fun i m ->
                Marshal.WriteInt32(pMaterials, 16 * i, 1)                                            // FT_MATERIAL_SOLID
                f32 pMaterials (16 * i + 4) m.R; f32 pMaterials (16 * i + 8) m.G; f32 pMaterials (16 * i + 12) m.B)
            lights |> Array.iteri (fun i l ->
                let kind, (vec:Vector3), (c:FColor) =
                    match l with
                    | Directional (d, c) -> 1, d, c                                                  // FT_LIGHT_DIRECTIONAL
                    | Point (p, c) -> 2, p, c                                                        // FT_LIGHT_POINT
                Marshal.WriteInt32(pLights, 28 * i, kind)
                [ vec.X; vec.Y; vec.Z; c.R; c.G; c.B ] |> List.iteri (fun k x -> f32 pLights (28 * i + 4 + 4 * k) x))
            let mutable desc : Native.SceneDesc =
                { forms = pForms; objects = pObjects; children = pChildren; materials = pMaterials; lights = pLights
                  nForms = forms.Count; nObjects = objects.Count; nChildren = children.Count
                  nMaterials = materials.Count; nLights = lights.Length; rootObject = root
                  bgR = scene.BackgroundColor.R; bgG = scene.BackgroundColor.G; bgB = scene.BackgroundColor.B }
            Native.ft_scene_create(&desc, device, &handle) |> Native.check      // the tables are copied: free them below
        finally
            for p in [ pForms; pObjects; pChildren; pMaterials; pLights ] do Marshal.FreeHGlobal p
    member _.Handle = handle
    interface IDisposable with
        member _.Dispose() = if handle <> 0n then Native.ft_scene_destroy handle |> ignore

/// A reusable FColor[X,Y] that is pinned AND registered with the CUDA driver (ft_host_register: page-locked and mapped into
/// the device's address space).  ft_render then takes its zero-copy path: the kernel stores every finished 4 x 8 pixel tile
/// straight into this array (128-bit stores over PCIe), no staging buffer and no device->host copy.  Registration costs a few
/// milliseconds once (bench.py prints `host_register_ms`), so keep one FrameBuffer per output size and render into it again.
type FrameBuffer(size:ImageSize) =
    let image = Array2D.zeroCreate<FColor> size.X size.Y
    let pin = GCHandle.Alloc(image, GCHandleType.Pinned)
    do Native.ft_host_register(pin.AddrOfPinnedObject(), int64 size.X * int64 size.Y * 12L) |> Native.check
    member _.Size = size
    member _.Image = image
    member _.Address = pin.AddrOfPinnedObject()
    interface IDisposable with
        member _.Dispose() =
            Native.ft_host_unregister(pin.AddrOfPinnedObject()) |> ignore
            pin.Free()

[<RequireQualifiedAccess>]
module Image =
    let private call (scene:CompiledScene) (camera:Camera) (size:ImageSize) (epsilon:float32) (length:float32) (address:nativeint) =
        let mutable cam : Native.CameraDesc =
            { position = camera.Position; forward = camera.Forward
              upScaled = camera.UpScaled; rightScaled = camera.RightScaled }
        Native.ft_render(scene.Handle, &cam, size.X, size.Y, epsilon, length, 0u, address, 0n, 0n) |> Native.check

    /// The fast path (bench.py `e2e`): renders into a registered FrameBuffer and returns its FColor[X,Y].
    let renderInto (frame:FrameBuffer) (epsilon:float32) (length:float32) (camera:Camera) (scene:CompiledScene) : FColor[,] =
        call scene camera frame.Size epsilon length frame.Address
        frame.Image

    /// Image.render epsilon length imageSize camera (SdfScene.trace scene)   (Image.fs:26-35), the reference's own signature:
    /// a FRESH FColor[X,Y] per call.  FColor[X,Y] is row-major with X the slow index and 12 bytes per element: exactly
    /// ft_render's out_rgb.  A fresh array is pinned but not registered (registering 25 MB costs more than one 1080p frame),
    /// so this call takes ft_render's staged path -- two column bands through device memory, bench.py `e2e_pageable`.
    /// Callers that render more than one frame should hold a FrameBuffer and use renderInto.
    let render (epsilon:float32) (length:float32) (imageSize:ImageSize) (camera:Camera) (scene:CompiledScene) : FColor[,] =
        let image = Array2D.zeroCreate<FColor> imageSize.X imageSize.Y
        let pin = GCHandle.Alloc(image, GCHandleType.Pinned)
        try call scene camera imageSize epsilon length (pin.AddrOfPinnedObject())
        finally pin.Free()
        image
